#!/usr/bin/env python
"""Regenerate profiles/traffic.json entries from an ncu launch list (read here, no GPU needed).

The launch list is the CSV of
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,\
sm__cycles_elapsed.avg,sm__inst_executed.avg.per_cycle_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed \
        --clock-control none -k regex:<solver kernels> \
        --csv --log-file <csv> python bench.py --steps K --warmup W --no-cpu --no-stream [--workload ...]
i.e. the SAME bench command whose line carries `roofline`; bench.py never measures under the profiler, it
only reads the per-step numbers this script writes:

    python tools/ncu_traffic.py <workload name> <launches.csv> [--kernels REGEX] [--out profiles/traffic.json]

Per step = totals over the matching launches / number of solver passes in the capture (= the most common
launch count among the matching kernels: every pass launches each size class once).  IPC of the pass = warp
instructions / (SM cycles x SMs), the cycle-weighted mean over its kernels (they are serialised under ncu);
l1_pipe = the L1/shared-memory data pipe (LSU wavefronts: LDS bank replays, LDG data, SHFL) as a fraction of its
peak, cycle-weighted likewise (absent from captures taken without that metric).
"""
import csv
import json
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_RE = r"tc_hash_kernel|tc_hybrid_kernel|tc_rank_kernel|tc_support_kernel|k_diamond_sum|k_motif4_closed|c4_\w+_kernel|kclique_bitmap_kernel|kclique_warp_edge"


def parse(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    for r in rd:
        if r[0] == "ID":
            continue
        lid = int(r[ix["ID"]])
        d = launches.setdefault(lid, {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]})
        try:
            d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
            d[r[ix["Metric Name"]] + ":unit"] = r[ix["Metric Unit"]]
        except ValueError:
            pass
    return [launches[k] for k in sorted(launches)]


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ns(v, unit):
    return v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}.get(unit, 1)


def short(name):
    m = re.search(r"(\w+)<([^>]*)>\(", name) or re.search(r"(\w+)\(", name)
    if not m:
        return name[:60]
    return m.group(1) + ("<" + m.group(2).replace("(int)", "").replace("(bool)", "").replace(" ", "") + ">" if m.lastindex == 2 else "")


def main():
    args = sys.argv[1:]
    if len(args) < 2:
        print(__doc__); sys.exit(2)
    name, path = args[0], args[1]
    kre = re.compile(args[args.index("--kernels") + 1] if "--kernels" in args else DEFAULT_RE)
    out_path = args[args.index("--out") + 1] if "--out" in args else os.path.join(ROOT, "profiles", "traffic.json")
    per = defaultdict(lambda: defaultdict(float))
    for l in parse(path):
        if not kre.search(l["name"]):
            continue
        k = short(l["name"])
        p = per[k]
        p["launches"] += 1
        p["dram_read"] += to_bytes(l.get("dram__bytes_read.sum", 0.0), l.get("dram__bytes_read.sum:unit", "byte"))
        p["dram_write"] += to_bytes(l.get("dram__bytes_write.sum", 0.0), l.get("dram__bytes_write.sum:unit", "byte"))
        p["warp_insts"] += l.get("smsp__inst_executed.sum", 0.0)
        p["sm_cycles"] += l.get("sm__cycles_elapsed.avg", 0.0)
        p["l1_pipe_cycles"] += l.get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 0.0) / 100.0 * l.get("sm__cycles_elapsed.avg", 0.0)
        p["time_ns"] += to_ns(l.get("gpu__time_duration.sum", 0.0), l.get("gpu__time_duration.sum:unit", "ns"))
        ipc = l.get("sm__inst_executed.avg.per_cycle_elapsed", 0.0)
        if ipc > 0 and l.get("sm__cycles_elapsed.avg", 0.0) > 0:
            p["sms_est"] = max(p["sms_est"], round(l["smsp__inst_executed.sum"] / (ipc * l["sm__cycles_elapsed.avg"])))
    if not per:
        sys.exit("no launch matches " + kre.pattern)
    # every pass launches each size class once; a tier that launches per root (c4_heavy_kernel) has many launches
    # per pass: the MOST COMMON launch count among the kernels is the number of passes
    counts = sorted(int(p["launches"]) for p in per.values())
    passes = max(set(counts), key=lambda c: (counts.count(c), -c))
    sms = max(int(p["sms_est"]) for p in per.values()) or 148
    tot = defaultdict(float)
    kernels = {}
    for k, p in sorted(per.items()):
        n = passes
        kernels[k] = {"launches_per_step": p["launches"] / n, "ms_per_step": p["time_ns"] / n / 1e6,
                      "dram_bytes_per_step": (p["dram_read"] + p["dram_write"]) / n,
                      "warp_insts_per_step": p["warp_insts"] / n,
                      "ipc": p["warp_insts"] / (p["sm_cycles"] * sms) if p["sm_cycles"] else None,
                      "l1_pipe": (p["l1_pipe_cycles"] / p["sm_cycles"]) if p["sm_cycles"] and p["l1_pipe_cycles"] else None}
        for f in ("dram_read", "dram_write", "warp_insts", "sm_cycles", "time_ns", "l1_pipe_cycles"):
            tot[f] += p[f] / n
    entry = {"dram_bytes_per_step": tot["dram_read"] + tot["dram_write"],
             "dram_read_per_step": tot["dram_read"], "dram_write_per_step": tot["dram_write"],
             "warp_insts_per_step": tot["warp_insts"], "ipc": tot["warp_insts"] / (tot["sm_cycles"] * sms) if tot["sm_cycles"] else None,
             "l1_pipe": (tot["l1_pipe_cycles"] / tot["sm_cycles"]) if tot["sm_cycles"] and tot["l1_pipe_cycles"] else None,
             "serialized_ms_per_step": tot["time_ns"] / 1e6, "passes_in_capture": passes, "sms": sms,
             "kernels": kernels, "source": os.path.relpath(os.path.abspath(path), ROOT) + " (tools/ncu_traffic.py)"}
    try:
        data = json.load(open(out_path))
    except Exception:
        data = {}
    data["_comment"] = ("per bench step, summed over the kernels of one solver pass; written by tools/ncu_traffic.py from the ncu launch "
                        "lists named in `source` (bench.py copies dram_bytes_per_step into roofline.traffic)")
    data[name] = entry
    json.dump(data, open(out_path, "w"), indent=1)
    print(json.dumps({name: {k: v for k, v in entry.items() if k != "kernels"}}, indent=1))
    for k, v in kernels.items():
        print(f"  {k:60s} {v['ms_per_step']:9.3f} ms  {v['dram_bytes_per_step'] / 1e9:8.3f} GB  {v['warp_insts_per_step'] / 1e9:8.3f} G inst  ipc {v['ipc'] or 0:.2f}  l1 pipe {v.get('l1_pipe') or 0:.2f}")


if __name__ == "__main__":
    main()
