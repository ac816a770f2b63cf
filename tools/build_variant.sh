#!/bin/bash
# Build libvgslm with experiment switches into a SEPARATE library, next to the product one, for A/B runs on the GPU box:
#   tools/build_variant.sh uniform "-DVG_GEMM_UNIFORM_ISSUE=1 -DVG_DL_UNIFORM_ISSUE=1"
#   VGSLM_LIB=$PWD/vae_gslm_b200/libvgslm_uniform.so python tools/gemm_bench.py --quick
# Objects go to csrc/build_<name>/ (git- and gpurun-ignored); the product library and its objects are not touched.
set -e
name=$1; flags=$2
cd "$(dirname "$0")/../vae_gslm_b200/csrc"
make -j8 BUILD=build_$name OUT=../libvgslm_$name.so EXTRA="$flags" 2>&1 | grep -v "^/usr/local/cuda/bin/nvcc" | tail -3
ls -la ../libvgslm_$name.so
