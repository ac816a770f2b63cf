# round 2, GPU call 1: measure what round 1 wrote but never ran.
#   expA = -DVG_GEMM_UNIFORM_ISSUE=1 -DVG_DL_UNIFORM_ISSUE=1 -DVG_ATTN_FWD_SPLIT=2
#   expB = -DVG_ATTN_BWD_PERSIST=1
# plus the decode-design microbenchmarks (grid barrier / same-address reduction / bulk-copy streaming rate).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
echo "== microbench"
timeout 180 python tools/microbench/run.py > gpurun_out/r2c1_microbench.log 2>&1; echo "rc=$?" >> gpurun_out/r2c1_microbench.log
tail -45 gpurun_out/r2c1_microbench.log
for name in expA expB; do
  lib=$PWD/vae_gslm_b200/libvgslm_$name.so
  echo "== parity of $name"
  VGSLM_LIB=$lib timeout 400 python -m pytest tests/test_kernels_gpu.py -q --tb=short -k "attention or gemm or decode_linear" > gpurun_out/r2c1_pytest_$name.log 2>&1; echo "rc=$?" >> gpurun_out/r2c1_pytest_$name.log
  tail -4 gpurun_out/r2c1_pytest_$name.log
done
for which in product expA expB; do
  if [ $which = product ]; then unset VGSLM_LIB; else export VGSLM_LIB=$PWD/vae_gslm_b200/libvgslm_$which.so; fi
  echo "== perf $which"
  timeout 120 python tools/attn_raw.py 2>&1 | tail -4 | tee gpurun_out/r2c1_attn_$which.log
  if [ $which != expB ]; then
    timeout 240 python tools/gemm_bench.py --quick 2>&1 | tail -22 | tee gpurun_out/r2c1_gemm_$which.log
    timeout 240 python tools/decode_bench.py 1 64 2>&1 | tail -3 | tee gpurun_out/r2c1_decode_$which.log
  fi
done
unset VGSLM_LIB
