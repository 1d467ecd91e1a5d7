#!/usr/bin/env python
"""Summarise an .ncu-rep of the analysis kernel: headline metrics, stall mix, opcode mix,
per-phase instruction / sample shares.  Usage: ncu_summary.py REPORT FRAMES [out.txt]"""
import collections
import csv
import io
import re
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, frames = sys.argv[1], float(sys.argv[2])
    out = io.StringIO()
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
            "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]
    print("== headline", file=out)
    for k in keys:
        if k in m:
            print(f"{k:70s} {m[k]:>18s} {u[k]}", file=out)
    try:
        inst = float(m["smsp__inst_executed.sum"])
        print(f"warp-instructions per frame: {inst / frames:.0f}", file=out)
    except Exception:
        pass
    print("== stall mix (warps stalled per issue-active cycle)", file=out)
    st = [(k, float(v)) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
    for k, v in sorted(st, key=lambda x: -x[1])[:10]:
        print(f"   {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:24s} {v:6.2f}", file=out)
    print("== pipe utilisation (pct of peak, active)", file=out)
    for k, v in sorted(m.items()):
        if re.match(r"sm__inst_executed_pipe_[a-z_]+\.sum\.pct_of_peak_sustained_active", k) and float(v) > 0.5:
            print(f"   {k.split('pipe_')[1].split('.')[0]:12s} {float(v):6.2f}", file=out)

    sass = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    h = sass[1]
    ci = {x: i for i, x in enumerate(h)}
    ops, samp = collections.Counter(), collections.Counter()
    for r in sass[2:]:
        if len(r) < 10:
            continue
        mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]].strip())
        if not mm:
            continue
        op = mm.group(2).split(".")[0]
        ops[op] += int(r[ci["Instructions Executed"]] or 0)
        samp[op] += int(r[ci["# Samples"]] or 0)
    tot, ts = sum(ops.values()), sum(samp.values())
    print(f"== opcode mix (warp-instr per frame; total {tot / frames:.0f})", file=out)
    for op, n in ops.most_common(24):
        print(f"   {op:10s} {n / frames:8.1f} {100 * n / tot:6.2f}%  samples {100 * samp[op] / ts:5.2f}%", file=out)

    cs = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur, data = None, []
    for r in cs:
        if r and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0].isdigit():
            try:
                data.append((cur, int(r[0]), r[1].strip(), int(r[4] or 0), int(r[7] or 0)))
            except Exception:
                pass
    toti, tots = sum(d[4] for d in data) or 1, sum(d[3] for d in data) or 1
    print("== hottest source lines by stall samples", file=out)
    for f, ln, src, s, i in sorted(data, key=lambda d: -d[3])[:25]:
        print(f"   {f}:{ln:<4d} samples {100 * s / tots:5.2f}%  inst {100 * i / toti:5.2f}%  {src[:90]}", file=out)
    print("== hottest source lines by instructions", file=out)
    for f, ln, src, s, i in sorted(data, key=lambda d: -d[4])[:25]:
        print(f"   {f}:{ln:<4d} inst {100 * i / toti:5.2f}%  samples {100 * s / tots:5.2f}%  {src[:90]}", file=out)
    text = out.getvalue()
    if len(sys.argv) > 3:
        open(sys.argv[3], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
