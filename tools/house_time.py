#!/usr/bin/env python
"""sgl house on the LiveJournal-shaped graph (optionally divided): the DAG fast path (supports + per-edge 4-cycle
counts) against the operator-API kernel (the reference's schedule).  python tools/house_time.py [div] [--list]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
div = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1
rp, ci = capi.generate_graph(4_847_571 // div, 68_993_773 // div, 0x5EED004C)
print("nv", rp.numel() - 1, "ne", ci.numel(), flush=True)
g = capi.DeviceGraph.adopt(rp, ci, 0); g.prepare("sgl:house")
for _ in range(3):
    c = g.sgl("house"); print("house (fast)", c, "%.2f ms" % g.last_stats()[0], g.last_stats()[1], "launches", flush=True)
if "--list" in sys.argv:
    capi.set_option("sgl.algo", "list")
    t0 = time.time(); c2 = g.sgl("house"); print("house (operator API)", c2, "%.2f ms" % g.last_stats()[0], "match", c2 == c, flush=True)
