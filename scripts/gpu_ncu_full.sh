#!/bin/bash
# one `ncu --set full --import-source on` capture of a kernel.  Usage: gpu_ncu_full.sh <tag> <kernel regex> <skip> <count> <python args...>
tag=$1; kre=$2; skip=$3; cnt=$4; shift 4
out=gpurun_out; mkdir -p $out
timeout 600 python "$@" > $out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $out/${tag}_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$kre" -s $skip -c $cnt -f -o $out/${tag} python "$@" > $out/${tag}_ncu.log 2>&1
tail -2 $out/${tag}_ncu.log
ls -la $out/${tag}.ncu-rep
