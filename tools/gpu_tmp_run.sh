tag=r2x
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${tag}_pytest.txt
for v in 0 1 0 1; do
MMDGPU_INLINE_IDS=$v timeout 300 python bench.py --workload C4 --instances 64 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-also > gpurun_out/${tag}_c4_64.json 2>gpurun_out/${tag}_err.txt
python - <<PY | tee -a gpurun_out/${tag}_ab.txt
import json
d=json.loads(open("gpurun_out/${tag}_c4_64.json").read().strip().splitlines()[-1])
print("INLINE_IDS=$v  C4 x 64 instances: %.1f G  %.1f us/step" % (d["value"]/1e9, d["ms_per_step"]*1e3))
PY
done
timeout 600 python tools/gpu_fuzz.py 3000 300 crowd 2>&1 | tail -2 | tee gpurun_out/${tag}_fuzz_crowd.txt
timeout 600 python tools/gpu_fuzz.py 3000 200 motion 2>&1 | tail -2 | tee gpurun_out/${tag}_fuzz_motion.txt
