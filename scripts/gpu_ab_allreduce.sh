#!/bin/bash
# A/B of the train step's exchange on N GPUs: the one-shot peer all-reduce against ncclAllReduce, on the graph-captured steps
# (batch 8192 per GPU, global batch 500) and the large-batch step.  Usage (under gpurun --gpus N): gpu_ab_allreduce.sh N [workloads]
n=$1; shift
wls=${*:-"sdss_train_b8192 sdss_train_b500"}
mkdir -p gpurun_out
for wl in $wls; do for ar in peer nccl peer nccl; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29542 bench.py \
    --gpus $n --steps 30 --warmup 5 --workload $wl --no-also --no-e2e --no-cpu-baseline --allreduce $ar > gpurun_out/ab_${n}_${wl}_$ar.log 2>&1
  grep "^{" gpurun_out/ab_${n}_${wl}_$ar.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$n GPUs', '$wl', '$ar', '->', d.get('config', {}).get('allreduce'), 'us/step %.1f' % (1e3 * d.get('ms_per_step', 0)), 'M/s %.2f' % (d.get('value', 0) / 1e6), d.get('dp_parity', {}).get('note'), d.get('error'))
"
done; done | tee gpurun_out/ab_allreduce_${n}gpu.txt
