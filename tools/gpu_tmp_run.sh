tag=r2u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-also --layout sokol32 > gpurun_out/${tag}_bench_sokol32.json 2>> gpurun_out/${tag}_bench.err
bash tools/ncu_capture.sh ${tag} skin_pair
python - <<PY
import json
for f in ("bench", "bench_sokol32", "ref"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.3f G" % (d["value"] / 1e9), "e2e", d.get("e2e", {}).get("value"), "frac", (d.get("roofline") or {}).get("frac"))
        for k, v in (d.get("also") or {}).items():
            print("   ", k, "%.2f G" % (v["value"] / 1e9), v.get("ms_per_step"))
    except Exception as e:
        print(f, "failed:", e)
PY
tail -3 gpurun_out/${tag}_bench.err
