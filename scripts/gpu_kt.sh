#!/bin/bash
# per-kernel times (ncu launch list, no clock control, only this library's kernels) of a python command.
# Usage: gpu_kt.sh <tag> <python args...>     e.g.  gpu_kt.sh l32 scripts/l32_time.py 65536
tag=$1; shift
out=gpurun_out; mkdir -p $out
timeout 600 python "$@" > $out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_plain.log; exit 1; }
tail -2 $out/${tag}_plain.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -k regex:'k_tc|k_gram|k_grad|k_reduce|k_adam|k_solve|k_gen|k_mma|k_out|k_gather|k_final|k_ood|k_sample' -c 200 --csv --log-file $out/${tag}_launches.csv python "$@" > $out/${tag}_ncu.log 2>&1
python scripts/launch_summary.py $out/${tag}_launches.csv | tee $out/${tag}_launches_summary.txt
