#!/bin/bash
# ncu launch list (gpu__time_duration per launch, no clock control) of one bench workload.  Usage: gpu_launchlist.sh <tag> <workload> [extra bench args]
tag=$1; wl=$2; shift 2
out=gpurun_out; mkdir -p $out
B="python bench.py --workload $wl --steps 3 --warmup 3 --no-also --no-e2e --no-cpu-baseline $*"
timeout 600 $B > $out/${tag}_${wl}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_${wl}_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -k regex:'k_tc|k_gram|k_grad|k_reduce|k_adam|k_solve|k_out|k_gather|k_final|k_ood' -c 120 --csv --log-file $out/${tag}_${wl}_launches.csv $B > $out/${tag}_${wl}_ncu.log 2>&1
python scripts/launch_summary.py $out/${tag}_${wl}_launches.csv | tee $out/${tag}_${wl}_launches_summary.txt
