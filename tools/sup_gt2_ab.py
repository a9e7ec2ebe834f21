#!/usr/bin/env python
"""A/B of the support kernel's class-2 group width on the LiveJournal-shaped graph (diamond)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import shaped_graph
rp, ci = shaped_graph(4_847_571, 68_993_773, 0x5EED004C, device="cuda:0")
for gt in (256, 512, 1024, 256, 512, 1024):
    capi.set_option("sup.gt2", gt)
    g = capi.DeviceGraph.adopt(rp, ci, 0); g.prepare("sgl:diamond")
    for _ in range(3): c = g.sgl("diamond")
    ts = sorted(g.sgl("diamond") and g.last_stats()[0] for _ in range(7))
    print(f"sup.gt2={gt}: median {ts[3]:.3f} ms best {ts[0]:.3f} count {c}", flush=True)
    g.close()
