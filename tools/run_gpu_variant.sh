# A/B of an experiment build against the product library on the GPU box (build it first, on the CPU side:
#   tools/build_variant.sh exp "-DVG_GEMM_UNIFORM_ISSUE=1 -DVG_DL_UNIFORM_ISSUE=1 -DVG_ATTN_FWD_SPLIT=2 -DVG_ATTN_BWD_PERSIST=1")
# usage: bash tools/run_gpu_variant.sh exp
name=${1:-exp}
lib=$PWD/vae_gslm_b200/libvgslm_$name.so
mkdir -p gpurun_out
echo "== parity of the variant (kernel tests through the C ABI)"
VGSLM_LIB=$lib timeout 300 python -m pytest tests/test_kernels_gpu.py -q --tb=short -x -k "attention or gemm or decode_linear" > gpurun_out/pytest_variant_$name.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_variant_$name.log
tail -3 gpurun_out/pytest_variant_$name.log
for which in product $name; do
  if [ $which = product ]; then unset VGSLM_LIB; else export VGSLM_LIB=$lib; fi
  echo "== $which"
  timeout 100 python tools/attn_raw.py 2>&1 | tail -4
  timeout 200 python tools/gemm_bench.py --quick 2>&1 | tail -22
  timeout 200 python tools/decode_bench.py 1 64 2>&1 | tail -4
done
unset VGSLM_LIB
timeout 120 python tools/microbench/run.py 2>&1 | tail -40
