tag=r3a
timeout 700 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${tag}_pytest.txt
for v in 1 0 1 0; do
MMDGPU_PADDED_SLOTS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/${tag}_b.json 2>gpurun_out/${tag}_err.txt
python - <<PY | tee -a gpurun_out/${tag}_ab.txt
import json
d=json.loads(open("gpurun_out/${tag}_b.json").read().strip().splitlines()[-1]); e=d["e2e"]
print("PADDED_SLOTS=$v  value %.1f G  e2e %.3f G  %.2f ms/step  d2h %.1f GB/s  ceiling %.1f  frac %.3f" % (d["value"]/1e9, e["value"]/1e9, e["ms_per_step"], e["d2h_gbs_per_gpu"], e["d2h_ceiling_gbs_per_gpu"], e["frac_of_d2h_ceiling"]))
PY
done
