#!/usr/bin/env python
"""Warp stall breakdown (pc-sampling) per kernel of an .ncu-rep: python tools/ncu_stalls.py report.ncu-rep [substr]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; filt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
cols = [(i, h[len("smsp__pcsamp_warps_issue_stalled_"):]) for i, h in enumerate(hdr)
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if filt and filt not in name: continue
    vals = [(float(r[i] or 0), n) for i, n in cols]; tot = sum(v for v, _ in vals) or 1
    print("== " + name[:70] + "  " + "  ".join(f"{n}={v / tot * 100:.0f}%" for v, n in sorted(vals, reverse=True) if v / tot > 0.03))
