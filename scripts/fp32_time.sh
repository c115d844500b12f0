#!/bin/bash
for w in sdss100k_predict sdss_train; do python bench.py --precision fp32 --workload $w --no-also --no-e2e --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'fp32', round(d['value']/1e6,2), 'M/s')"; done | paste - -
