mkdir -p gpurun_out
for mode in gelu; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_$mode \
  python tools/gemm_one.py 8000 4096 1024 $mode > gpurun_out/ncu_gemm_$mode.log 2>&1; tail -2 gpurun_out/ncu_gemm_$mode.log
done
