#!/bin/bash
# device-resident bench of every workload, one summary line each
for w in C3 C4 C1 C2; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also "$@" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernel_ms_per_step']; print(d['config']['workload'][:3], 'value %.2f G  ms/step %.4f  skin %.4f hier %.4f K1 %.4f frac %.3f'%(d['value']/1e9,d['ms_per_step'],k['skin'],k['hierarchy'],k['pose_sample'],d['roofline']['frac']))
    else: print(l.rstrip())
"
done
