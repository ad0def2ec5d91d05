#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch) into a small JSON for profiles/: run here, no GPU needed.
usage: summarize_ncu.py <rep> <out.json> [placements] [ops_per_placement]"""
import csv, json, subprocess, sys, collections

rep, out = sys.argv[1], sys.argv[2]
placements = float(sys.argv[3]) if len(sys.argv) > 3 else None
ops = float(sys.argv[4]) if len(sys.argv) > 4 else 220.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, row = rows[0], rows[1], rows[2]
d = {h: (row[i], units[i]) for i, h in enumerate(hdr)}
extra_launches = rows[3:]               # a report with several launches (e.g. the re-spreading stages of one play): summarised below
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max"]
res = {"report": rep.split("/")[-1], "kernel": d.get("Kernel Name", ("?",))[0], "metrics": {}}
for k in keys:
    if k in d:
        v, u = d[k]
        try:
            v = float(v)
        except ValueError:
            pass
        res["metrics"][k] = {"value": v, "unit": u}
def to_bytes(k):
    v, u = d.get(k, ("0", "byte"))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return float(v) * mult
res["dram_bytes_per_launch"] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
if placements:
    res["placements_per_launch"] = placements
    res["dram_bytes_per_placement"] = res["dram_bytes_per_launch"] / placements
    res["warp_instructions_per_placement"] = float(d["smsp__inst_executed.sum"][0]) / placements
    res["algorithmic_thread_ops_per_placement"] = ops
# stall reasons from the source page
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
h = None
tot = collections.Counter()
for r in srows:
    if r and r[0] == "Address":
        h = r
        continue
    if h is None or not r or not r[0].startswith("0x"):
        continue
    for i, name in enumerate(h):
        if name.startswith("stall_") and "Not Issued" not in name and i < len(r):
            try:
                tot[name] += int(r[i] or 0)
            except ValueError:
                pass
s = sum(tot.values()) or 1
res["warp_stall_samples_pct"] = {k: round(100.0 * v / s, 2) for k, v in tot.most_common(8)}
if extra_launches:
    res["launches"] = []
    for r in [row] + extra_launches:
        dd = {h: r[i] for i, h in enumerate(hdr) if i < len(r)}
        one = {"kernel": dd.get("Kernel Name", "?")}
        for k in keys:
            if k in dd:
                try:
                    one[k] = float(dd[k])
                except ValueError:
                    one[k] = dd[k]
        one["unit_of"] = {k: d[k][1] for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum") if k in d}
        res["launches"].append(one)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1)[:1500])
