#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list, ncu full captures of the hot kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.csv
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$B > $out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_tc|k_gram|k_grad|k_reduce|k_adam|k_final|k_clip|k_smooth|k_prepare|k_solve' -c 400 --csv --log-file $out/${tag}_launches.csv $B > $out/${tag}_ncu_launch.log 2>&1
P="python scratch/tc_prof.py 17760 predict"
$P > $out/${tag}_plain_p.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tc_gram -s 1 -c 1 -f -o $out/${tag}_predict $P > $out/${tag}_ncu_p.log 2>&1
T="python scratch/tc_prof.py 17760 train"
$T > $out/${tag}_plain_t.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_tc_gram|k_tc_grad|k_reduce' -s 3 -c 3 -f -o $out/${tag}_train $T > $out/${tag}_ncu_t.log 2>&1
L="python scratch/l32_time.py 17760"
$L > $out/${tag}_plain_l.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_tc_gram32|k_solve32|k_tc_grad32' -s 9 -c 3 -f -o $out/${tag}_l32 $L > $out/${tag}_ncu_l.log 2>&1
python bench.py --workload desi_score --no-also --no-e2e --steps 5 > $out/${tag}_bench_desi.json 2>> $out/${tag}_bench.err
cat $out/${tag}_bench.json
