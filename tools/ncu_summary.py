#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-kernel headline metrics + hottest SASS lines.
usage: python tools/ncu_summary.py report.ncu-rep [kernel-substring] [--sass N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
nsass = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if filt and filt not in name: continue
    print("== " + name[:110])
    for w in WANT:
        if w in idx: print(f"   {w:62s} {r[idx[w]]} {units[idx[w]]}")
if nsass:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    secs, cur = [], None
    for r in csv.reader(io.StringIO(src)):
        if len(r) >= 2 and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; secs.append(cur); continue
        if r and r[0] == "Address": cur["hdr"] = r; continue
        if cur is not None and r: cur["rows"].append(r)
    seen = set()
    for s in secs:
        if (filt and filt not in s["name"]) or s["name"] in seen: continue
        seen.add(s["name"]); h = {n: i for i, n in enumerate(s["hdr"])}
        tot = sum(int(r[h["Instructions Executed"]] or 0) for r in s["rows"]) or 1
        ts = sum(int(r[h["# Samples"]] or 0) for r in s["rows"]) or 1
        print(f"-- SASS {s['name'][:100]}  inst={tot} samples={ts}")
        top = sorted(s["rows"], key=lambda r: -int(r[h["# Samples"]] or 0))[:nsass]
        keep = {id(r) for r in top}
        for r in s["rows"]:
            if id(r) in keep:
                print(f"   inst {int(r[h['Instructions Executed']] or 0) / tot * 100:5.2f}%  stall-samples {int(r[h['# Samples']] or 0) / ts * 100:5.2f}%  thr {r[h['Avg. Threads Executed']]:>4}  {r[h['Source']][:90]}")
