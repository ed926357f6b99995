#!/bin/bash
# scatter sub-range / L2 fetch granularity sweep at k=24 (stage timers from bench.py)
for sb in 0 1 2 3 4; do
  B200ZK_MSM_SUB_BITS=$sb python bench.py --k 24 --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('sub_bits=$sb', 'value', round(d['value'],1), {k:round(v,2) for k,v in d['msm']['stages_ms'].items()}, 'unreg', round(d['msm_unregistered']['ms'],2))
"
done
for g in 32 64 128; do
  B200ZK_L2_FETCH=$g python bench.py --k 24 --steps 2 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('l2fetch=$g', 'value', round(d['value'],1), {k:round(v,2) for k,v in d['msm']['stages_ms'].items()}, 'ntt', round(d['ntt']['ms'],3))
"
done
