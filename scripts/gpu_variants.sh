#!/bin/bash
# A/B of compile-time library variants (scripts/build_variants.py) on the GPU box:
#   bash scripts/gpu_variants.sh <name prefix> <command ...>
# every libsafeincave_cuda_<prefix>*.so is copied over the product's library and the command is run (then the default one).
cd "$(dirname "$0")/.."
PREFIX=$1; shift
CMD="$@"
cp safeincave_b200/libsafeincave_cuda.so /tmp/libsafeincave_cuda.default.so
for lib in safeincave_b200/libsafeincave_cuda_${PREFIX}*.so; do
  name=$(basename $lib .so); name=${name#libsafeincave_cuda_}
  cp $lib safeincave_b200/libsafeincave_cuda.so
  echo "== variant $name"
  timeout 300 $CMD 2>&1 | grep -h "CONSTITUTIVE\|ms/step\|OPERATOR\|rror" | tail -n 4
done
cp /tmp/libsafeincave_cuda.default.so safeincave_b200/libsafeincave_cuda.so
echo "== default build"
timeout 300 $CMD 2>&1 | grep -h "CONSTITUTIVE\|ms/step\|OPERATOR\|rror" | tail -n 4
