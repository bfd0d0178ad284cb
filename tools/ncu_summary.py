#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU): key raw metrics per captured launch and the SASS hot spots.
usage: tools/ncu_summary.py <file.ncu-rep> [--top N]"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
       "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
       "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
       "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:60])
        for w in RAW:
            if w in hdr:
                print(f"  {w:70s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
    src = run([rep, "--page", "source", "--csv"])
    blocks = src.split('"Kernel Name",')
    if len(blocks) < 2:
        return
    rows = list(csv.reader(io.StringIO(blocks[1].split("\n", 1)[1])))
    h = rows[0]
    iS, iI, iSamp, iT = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
    ins = [(r[iS].strip(), int(r[iI] or 0), int(r[iSamp] or 0), int(r[iT] or 0)) for r in rows[1:] if len(r) > iT and r[iI].isdigit()]
    tot_i, tot_s = sum(x[1] for x in ins), sum(x[2] for x in ins)
    print(f"-- first launch: {tot_i} warp instructions, {tot_s} stall samples, {len(ins)} SASS lines")
    by_op = collections.Counter()
    by_op_s = collections.Counter()
    for s, i, sm, _ in ins:
        op = s.split()[0] if not s.startswith("@") else s.split()[1]
        op = op.split(".")[0]
        by_op[op] += i
        by_op_s[op] += sm
    print("-- warp instructions by opcode (share of issue, share of stall samples)")
    for op, c in by_op.most_common(18):
        print(f"  {op:10s} {100.0 * c / tot_i:5.1f}%   samples {100.0 * by_op_s[op] / max(1, tot_s):5.1f}%")
    print(f"-- top {top} SASS lines by stall samples")
    for k, (s, i, sm, _) in sorted(enumerate(ins), key=lambda kv: -kv[1][2])[:top]:
        print(f"  #{k:4d} {100.0 * sm / max(1, tot_s):5.1f}%  exec {i:9d}  {s[:90]}")


if __name__ == "__main__":
    main()
