#!/usr/bin/env python
"""Summarise gpurun_out/ ncu artefacts into profiles/ (tracked).  Needs no GPU.

    python tools/ncu_summary.py r02

reads  gpurun_out/launches_<tag>.csv                      (ncu --metrics gpu__time_duration.sum launch list of `python bench.py`)
       gpurun_out/prof_<tag>_<workload>.ncu-rep           (ncu --set full captures; <workload> = c2, c5share, c3, c4, c2single, varint ...)
writes profiles/<tag>_ncu_summary.json                    (machine readable; bench.py reads `roofline.traffic` from it, by workload)
       profiles/<tag>_launches.md                         (the same as tables)
"""
import collections
import csv
import glob
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "profiles")
SRC = os.path.join(REPO, "gpurun_out")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size",
           "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg", "sm__cycles_elapsed.max",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(tag):
    path = os.path.join(SRC, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return []
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") == "gpu__time_duration.sum":
            agg[row["Kernel Name"].split("(")[0]].append(float(row["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    out = []
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        v2 = sorted(v)
        out.append({"kernel": k, "launches": len(v), "min_ns": v2[0], "median_ns": v2[len(v2) // 2], "max_ns": v2[-1], "sum_ns": sum(v),
                    "share_of_gpu_time": sum(v) / total})
    return out


def full(tag):
    """one record per captured launch, from the reports (or from their `--page raw --csv` exports made on the GPU box: the reports
    themselves are 4-5 MB per launch and gpurun brings back 64 MiB at most)"""
    out = []
    names = sorted(set(glob.glob(os.path.join(SRC, f"prof_{tag}_*.ncu-rep")) + glob.glob(os.path.join(SRC, f"prof_{tag}_*.raw.csv"))))
    seen = set()
    for rep in names:
        workload = os.path.basename(rep)[len(f"prof_{tag}_"):].replace(".ncu-rep", "").replace(".raw.csv", "")
        if workload in seen:
            continue
        seen.add(workload)
        if rep.endswith(".raw.csv"):
            raw = open(rep).read()
        else:
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rows = list(csv.reader([ln for ln in raw.splitlines() if not ln.startswith("==")]))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            rec = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "workload": workload, "report": os.path.basename(rep)}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    rec[m] = {"value": r[i], "unit": units[i]}
            out.append(rec)
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(OUT, exist_ok=True)
    L, F = launches(tag), full(tag)
    how = {}
    hp = os.path.join(SRC, f"how_{tag}.json")
    if os.path.exists(hp):
        how = json.load(open(hp))
    with open(os.path.join(OUT, f"{tag}_ncu_summary.json"), "w") as fh:
        json.dump({"launch_list": L, "full_capture": F, "how": how}, fh, indent=1)
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as fh:
        fh.write(f"# ncu launch list, {tag} (cold-cache, serialised: compare shares, not absolutes)\n\n")
        if how.get("launch_list"):
            fh.write(f"`{how['launch_list']}`\n\n")
        fh.write("| kernel | launches | min ns | median ns | max ns | share of GPU time |\n|---|---|---|---|---|---|\n")
        for r in L:
            fh.write(f"| {r['kernel']} | {r['launches']} | {r['min_ns']:.0f} | {r['median_ns']:.0f} | {r['max_ns']:.0f} | {r['share_of_gpu_time']:.3f} |\n")
        fh.write("\n# ncu --set full, per captured launch\n\n")
        for r in F:
            fh.write(f"## {r['kernel']}  ({r['workload']}, {r['report']})\n\n")
            for m in METRICS:
                if m in r:
                    fh.write(f"- `{m}` = {r[m]['value']} {r[m]['unit']}\n")
            fh.write("\n")
    print(open(os.path.join(OUT, f"{tag}_launches.md")).read()[:3000])


if __name__ == "__main__":
    main()
