#!/bin/bash
# GPU box: K3 iteration loop - the skinning-heavy parity tests, then the device-resident headline in both layouts + C1/C4.
tag=${1:-k3}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${tag}_pytest.txt
cat gpurun_out/${tag}_pytest.txt
for args in "" "--layout sokol32" "--workload C4" "--workload C1 --frames-per-step 512"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also $args 2>>gpurun_out/${tag}_err.txt | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$args'.ljust(40), 'value %.2f G  ms/step %.4f  skin %.4f ms  frac %.3f' % (d['value']/1e9, d['ms_per_step'], d['kernel_ms']['skin_per_launch_in_step'], d['roofline']['frac']))
" | tee -a gpurun_out/${tag}_bench.txt
done
