#!/bin/bash
# A/B of compile-time variants on the GPU box: ab_variants.sh "<python cmd>" "<nvcc extra 1>" "<nvcc extra 2>" ...
cmd=$1; shift
cp qfa_b200/libqfa_b200.so /tmp/lib_keep.so
for v in "$@"; do
  echo "=== variant: $v"
  QFA_NVCC_EXTRA="$v" python -c "from qfa_b200 import _lib; _lib.build(force=True)" 2>&1 | tail -3
  for i in 1 2; do timeout 300 python $cmd 2>&1 | tail -1; done
done
cp /tmp/lib_keep.so qfa_b200/libqfa_b200.so
