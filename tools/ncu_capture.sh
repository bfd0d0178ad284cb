#!/bin/bash
# Runs on the GPU box (under gpurun): plain run, then the ncu launch list and one full capture per kernel regex.
# usage: tools/ncu_capture.sh <tag> <kernel-regex> [extra bench args]
set -u
tag=$1; shift
kre=$1; shift
cmd="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-also $*"
mkdir -p gpurun_out
$cmd > gpurun_out/${tag}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu_list.log 2>&1
$cmd > gpurun_out/${tag}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${kre} -s 3 -c 2 -f -o gpurun_out/${tag}_full $cmd > gpurun_out/${tag}_ncu_full.log 2>&1
tail -n 2 gpurun_out/${tag}_ncu_full.log
