tag=r2s
line() {
  python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1'.ljust(60), 'value %.2f G  ms/step %.4f  skin %.4f ms' % (d['value']/1e9, d['ms_per_step'], d['kernel_ms']['skin_per_launch_in_step']))
" | tee -a gpurun_out/${tag}_ab.txt
}
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also"
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.txt
timeout 300 python tools/latency_probe.py 2>&1 | tee gpurun_out/${tag}_latency.txt
for sp in 0 1; do
  MMDGPU_IK_SPLIT=$sp $B --workload C2 --frames-per-step 128 2>>gpurun_out/${tag}_err.txt | line "IK_SPLIT=$sp C2 x 128"
  MMDGPU_IK_SPLIT=$sp $B --workload C2 --frames-per-step 256 2>>gpurun_out/${tag}_err.txt | line "IK_SPLIT=$sp C2 x 256"
done
