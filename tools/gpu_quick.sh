#!/bin/bash
# quick GPU check: parity tests + short device-resident bench; prints a one-line summary
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also "$@" | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['kernel_ms_per_step']; print('value %.2f G  ms/step %.3f  skin %.4f ms  hier %.3f  K1 %.3f frac %.3f'%(d['value']/1e9,d['ms_per_step'],k['skin'],k['hierarchy'],k['pose_sample'],d['roofline']['frac']))
    else: print(l.rstrip())
"
