mkdir -p gpurun_out
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 2 -o gpurun_out/prof_gemm_ffn1 -f python tools/gemm_one.py 8000 4096 1024 > gpurun_out/ncu_gemm.log 2>&1
cat gpurun_out/gemm_bench.log; tail -3 gpurun_out/ncu_gemm.log
