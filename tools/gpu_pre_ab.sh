#!/bin/bash
# GPU box: A/B of the fused-update pipelining depth (MMDGPU_PRE_STREAMS) on the IK model at small batches and on the headline,
# and of the staged / direct sokol32 output.  usage: tools/gpu_pre_ab.sh <tag>
tag=${1:-pre}
mkdir -p gpurun_out
line() {
  python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1'.ljust(60), 'value %.2f G  ms/step %.4f  skin %.4f ms' % (d['value']/1e9, d['ms_per_step'], d['kernel_ms']['skin_per_launch_in_step']))
" | tee -a gpurun_out/${tag}_ab.txt
}
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also"
for pre in 2 4; do
  for n in 128 256 512; do
    MMDGPU_PRE_STREAMS=$pre $B --workload C2 --frames-per-step $n 2>>gpurun_out/${tag}_err.txt | line "PRE=$pre C2 x $n"
  done
  MMDGPU_PRE_STREAMS=$pre $B 2>>gpurun_out/${tag}_err.txt | line "PRE=$pre C3"
  MMDGPU_PRE_STREAMS=$pre $B --workload C4 2>>gpurun_out/${tag}_err.txt | line "PRE=$pre C4"
done
for s in 0 1; do
  MMDGPU_SOKOL_STAGED=$s $B --layout sokol32 2>>gpurun_out/${tag}_err.txt | line "SOKOL_STAGED=$s C3 sokol32"
  MMDGPU_SOKOL_STAGED=$s $B --layout sokol32 --workload C4 2>>gpurun_out/${tag}_err.txt | line "SOKOL_STAGED=$s C4 sokol32"
done
