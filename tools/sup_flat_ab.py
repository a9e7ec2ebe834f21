#!/usr/bin/env python
"""A/B of the support pass's stream loop (sup.flat = 0: a loop per partner record | 1: flat windows) on the
LiveJournal-shaped graph (diamond) and, with an argument, the Friendster shape / N (formula 4-motif)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import shaped_graph
rp, ci = shaped_graph(4_847_571, 68_993_773, 0x5EED004C, device="cuda:0")
counts = set()
for flat in (0, 1, 0, 1):
    capi.set_option("sup.flat", flat)
    g = capi.DeviceGraph.adopt(rp, ci, 0); g.prepare("sgl:diamond")
    for _ in range(3): c = g.sgl("diamond")
    ts = sorted(g.sgl("diamond") and g.last_stats()[0] for _ in range(7))
    counts.add(c)
    print(f"sup.flat={flat}: median {ts[3]:.3f} ms best {ts[0]:.3f} count {c}", flush=True)
    g.close()
assert len(counts) == 1, counts
