"""Summarise ncu artefacts from gpurun_out/ into profiles/ (tracked).  Usage:
    python scripts/summarize_ncu.py <tag> [--rep gpurun_out/prof_pair.ncu-rep] [--launches gpurun_out/launches.csv]
Writes profiles/<tag>_launches.csv (per-kernel aggregate of the launch list), profiles/<tag>_<kernel>_metrics.csv (the
raw-page metrics that matter for a SIMT FP32 kernel), profiles/<tag>_<kernel>_stalls.txt (stall sampling by
instruction) and updates profiles/traffic.json (DRAM bytes per launch, read by bench.py)."""
import argparse
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmalite_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def launches(path, tag):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    ix = {h: i for i, h in enumerate(rows[hi])}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(ix) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v, u = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
        us = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3}.get(u, v)
        a = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    dst = os.path.join(ROOT, "profiles", f"{tag}_launches.csv")
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("kernel,launches,total_us,avg_us,share_pct\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.1f},{t / n:.1f},{100 * t / tot:.2f}\n")
    print("wrote", dst)


def kernel(rep, tag):
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    ix = {h: i for i, h in enumerate(hdr)}
    traffic = {}
    for row in raw[2:]:
        name = row[ix["Kernel Name"]].split("(")[0].split("::")[-1].split("<")[0].split()[-1]
        dst = os.path.join(ROOT, "profiles", f"{tag}_{name}_metrics.csv")
        with open(dst, "w") as f:
            f.write(f"# ncu --set full --clock-control none; kernel {row[ix['Kernel Name']]}\nmetric,unit,value\n")
            for k in KEEP:
                if k in ix:
                    f.write(f"{k},{units[ix[k]]},{row[ix[k]]}\n")
        print("wrote", dst)

        def val(k):
            v, u = float(row[ix[k]]), units[ix[k]].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

        traffic[f"{name}_dram_bytes_per_launch"] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    old.update(traffic)
    old["source"] = f"{tag}: ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch"
    json.dump(old, open(tpath, "w"), indent=1)
    src = ncu_csv(rep, "source")
    # the source page holds one block per profiled launch: a "Kernel Name" row, a header row, then instructions
    blocks, cur = [], None
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    seen = set()
    for blk in blocks:
        name = blk["name"].split("(")[0].split("::")[-1].split("<")[0].split()[-1]
        if name in seen:
            continue
        seen.add(name)
        h, data = blk["hdr"], blk["rows"]
        ix = {x: i for i, x in enumerate(h)}
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        kinds = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
        dst = os.path.join(ROOT, "profiles", f"{tag}_{name}_stalls.txt")
        with open(dst, "w") as f:
            f.write(f"# {blk['name']}\n# warp stall sampling, {tot} samples; share by reason\n")
            for k in sorted(kinds, key=lambda k: -sum(int(r[ix[k]]) for r in data)):
                n = sum(int(r[ix[k]]) for r in data)
                if n:
                    f.write(f"{k:28s} {100 * n / tot:6.2f}%\n")
            f.write("\n# top 40 instructions by samples\n")
            for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:40]:
                st = {k[6:]: int(r[ix[k]]) for k in kinds if int(r[ix[k]])}
                top = ", ".join(f"{k}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
                f.write(f"{int(r[ix['# Samples']]):7d}  {r[ix['Source']].strip()[:64]:64s} {top}\n")
        print("wrote", dst)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--rep")
    ap.add_argument("--launches")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    if a.launches:
        launches(a.launches, a.tag)
    if a.rep:
        kernel(a.rep, a.tag)
