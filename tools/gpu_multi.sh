#!/bin/bash
# GPU box with N GPUs: shard equivalence, D2H ceiling, and the driver's bench command at N.   usage: tools/gpu_multi.sh <N> <tag>
N=$1; tag=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1; nproc >> gpurun_out/${tag}_topo.txt
timeout 600 $TR 29511 tools/shard_equivalence.py > gpurun_out/${tag}_shard_equivalence.txt 2>gpurun_out/${tag}_shard_equivalence.err; cat gpurun_out/${tag}_shard_equivalence.txt
timeout 300 $TR 29512 tools/d2h_ceiling.py > gpurun_out/${tag}_d2h.json 2>gpurun_out/${tag}_d2h.err; cat gpurun_out/${tag}_d2h.json
timeout 300 $TR 29513 tools/d2h_ceiling.py --bind-numa > gpurun_out/${tag}_d2h_numa.json 2>>gpurun_out/${tag}_d2h.err; cat gpurun_out/${tag}_d2h_numa.json
timeout 1200 $TR 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2>gpurun_out/${tag}_bench.err
tail -c 1500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
