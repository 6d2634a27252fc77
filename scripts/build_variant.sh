#!/bin/bash
# Build the library with extra nvcc flags into build_variants/<name>.so (A/B runs inside one gpurun call:
# PYGLM_B200_LIB=build_variants/<name>.so python scripts/fused_probe.py), then restore the default build.
#   scripts/build_variant.sh epi1 -DPYGLM_TC_EPI=1
set -e
name=$1; shift
mkdir -p build_variants
PYGLM_NVCC_EXTRA="$*" python -c "from theano_pyglm_b200 import build; build.build(force=True)" > /dev/null
cp theano_pyglm_b200/lib/libpyglm_b200.so build_variants/$name.so
python -c "from theano_pyglm_b200 import build; build.build(force=True)" > /dev/null
echo built build_variants/$name.so
