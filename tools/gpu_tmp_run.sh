tag=r2t
line() {
  python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1'.ljust(60), 'value %.2f G  ms/step %.4f  skin %.4f ms' % (d['value']/1e9, d['ms_per_step'], d['kernel_ms']['skin_per_launch_in_step']))
" | tee -a gpurun_out/${tag}_ab.txt
}
B="timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also"
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${tag}_pytest.txt
timeout 300 python tools/latency_probe.py 2>&1 | tee gpurun_out/${tag}_latency.txt
timeout 200 python tools/latency_breakdown.py 2>&1 | tee gpurun_out/${tag}_latency_breakdown.txt
$B 2>>gpurun_out/${tag}_err.txt | line "C3"
$B --workload C4 2>>gpurun_out/${tag}_err.txt | line "C4"
