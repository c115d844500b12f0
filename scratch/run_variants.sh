#!/bin/bash
# A/B several builds of the library on the SAME box: scratch/run_variants.sh <script> [args]
for rep in 1 2; do
for so in scratch/lib_*.so; do
  cp $so qfa_b200/libqfa_b200.so
  echo "=== $so (rep $rep)"
  timeout 200 python "$@" 2>&1 | grep -E "mixed|accumulate|l32|Error|error" | head -12
done
done
