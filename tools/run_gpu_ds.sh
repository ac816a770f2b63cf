mkdir -p gpurun_out
bash tools/run_gpu_ncu_gemm.sh
timeout 200 ncu --set full --clock-control none --import-source on -k regex:skinny_linear -s 20 -c 1 -f -o gpurun_out/prof_skinny_b64 python tools/skinny_bench.py 64 > /dev/null 2>&1; ls -la gpurun_out/prof_skinny_b64.ncu-rep
