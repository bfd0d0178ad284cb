#!/bin/bash
# GPU box: full parity suite, then a short headline bench (device-resident only) and the complete default bench line.
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also > gpurun_out/${tag}_quick.json 2> gpurun_out/${tag}_quick.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-also --layout sokol32 > gpurun_out/${tag}_quick_sokol32.json 2>> gpurun_out/${tag}_quick.err
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_pytest.txt
python - <<PY
import json
for f in ("quick", "quick_sokol32", "bench"):
    try:
        d = json.loads(open("gpurun_out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.2f G" % (d["value"] / 1e9), "ms/step %.4f" % d["ms_per_step"], "skin %.4f" % d["kernel_ms"]["skin_per_launch_in_step"], "frac %.3f" % d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed:", e)
PY
tail -5 gpurun_out/${tag}_bench.err
