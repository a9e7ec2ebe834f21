#!/usr/bin/env python
"""sgl rectangle (all 4-cycles) on the LiveJournal-shaped graph through the wedge-pair fast path."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import shaped_graph
rp, ci = shaped_graph(4_847_571, 68_993_773, 0x5EED004C, device="cuda:0")
g = capi.DeviceGraph.adopt(rp, ci, 0); g.prepare("sgl:rectangle")
for _ in range(3):
    c = g.sgl("rectangle"); print("rectangle", c, "%.2f ms" % g.last_stats()[0], g.last_stats()[1], "launches", flush=True)
