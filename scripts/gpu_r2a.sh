#!/bin/bash
# GPU visit A of round 2: parity tests, then the default bench (both arms).  Usage under gpurun: bash scripts/gpu_r2a.sh <tag>
tag=${1:-r2a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.csv
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -15 $out/${tag}_pytest_gpu.log
( time timeout 900 python bench.py ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
tail -5 $out/${tag}_bench.err
cat $out/${tag}_bench.json | head -c 6000
