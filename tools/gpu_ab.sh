#!/bin/bash
# GPU box: A/B of an environment knob on the device-resident headline.  usage: tools/gpu_ab.sh <tag> <VAR> <values...>
tag=$1; var=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  for args in "" "--layout sokol32" "--workload C4"; do
    env $var=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also $args 2>>gpurun_out/${tag}_err.txt | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$var=$v', '$args'.ljust(20), 'value %.2f G  skin %.4f ms' % (d['value']/1e9, d['kernel_ms']['skin_per_launch_in_step']))
" | tee -a gpurun_out/${tag}_ab.txt
  done
done
