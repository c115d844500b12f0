#!/bin/bash
# Round-2 evidence run on one B200: tests, both bench arms, launch lists, ncu --set full captures.  bash scripts/gpu_r2_final.sh <tag>
tag=${1:-r2}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.csv
( time timeout 1500 python -m pytest tests -m gpu -q ) > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log
( time timeout 900 python bench.py ) > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
( time timeout 900 python bench.py --impl reference ) > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err; echo "ref rc=$?"
# LIGHT=1: launch lists only for the workloads whose kernels changed since the last full visit, no ncu --set full captures
WL="sdss100k_predict sdss_train l32_train desi_score l32_predict sdss_train_b8192 sdss_train_b500 sdss_train_tf32x3 sdss100k_predict_tf32x3"
[ -n "$LIGHT" ] && WL="sdss100k_predict l32_train sdss_train_b8192 sdss_train_b500"
rm -f $out/${tag}_launches_summary.txt
for w in $WL; do
  B="python bench.py --workload $w --steps 3 --warmup 3 --no-also --no-e2e --no-cpu-baseline"
  timeout 600 $B > $out/${tag}_${w}_plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -k regex:'k_tc|k_gram|k_grad|k_reduce|k_adam|k_solve|k_out|k_gather|k_final|k_ood' -c 120 --csv --log-file $out/${tag}_${w}_launches.csv $B > $out/${tag}_${w}_ncu.log 2>&1
  echo "## $w" >> $out/${tag}_launches_summary.txt
  python scripts/launch_summary.py $out/${tag}_${w}_launches.csv >> $out/${tag}_launches_summary.txt
done
cat $out/${tag}_launches_summary.txt
[ -n "$LIGHT" ] && exit 0
P="python scripts/tc_prof.py 17760 predict"
$P > $out/${tag}_plain_p.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gram -s 1 -c 1 -f -o $out/${tag}_predict $P > $out/${tag}_ncu_p.log 2>&1
T="python scripts/tc_prof.py 17760 train"
$T > $out/${tag}_plain_t.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tc_gram|k_tc_grad|k_reduce' -s 3 -c 3 -f -o $out/${tag}_train $T > $out/${tag}_ncu_t.log 2>&1
L="python scripts/l32_time.py 17760"
$L > $out/${tag}_plain_l.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tc_gram32|k_solve32|k_tc_grad32|k_reduce' -s 12 -c 4 -f -o $out/${tag}_l32 $L > $out/${tag}_ncu_l.log 2>&1
ls -la $out/*.ncu-rep
