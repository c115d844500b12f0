#!/bin/bash
for so in scratch/lib_*.so; do
  cp $so qfa_b200/libqfa_b200.so
  echo "=== $so"
  timeout 200 python "$@" 2>&1 | grep -E "mixed|Error|error" | head -12
done
