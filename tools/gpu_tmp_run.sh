tag=r3b_n8
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 900 $TR 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2>gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${tag}_bench.json").read().strip().splitlines() if l.startswith("{")][-1])
print("value %.1f G" % (d["value"] / 1e9), "e2e %.3f G" % (d["e2e"]["value"]/1e9), "frac_ceiling %.3f" % d["e2e"]["frac_of_d2h_ceiling"], "frac", d["roofline"]["frac"])
for k, v in (d.get("also") or {}).items():
    print("   ", k, "%.2f G" % (v["value"] / 1e9), v.get("ms_per_step"))
c5 = d["also"]["C5"]; print(json.dumps(c5.get("gather"))[:400]); print(json.dumps(c5.get("p2p_fused"))[:300])
PY
tail -3 gpurun_out/${tag}_bench.err
