import sys, os
sys.path.insert(0, "/root/repo")
import torch, bench
from graphminer_b200 import capi
for scale in (22, 24):
    rp, ci = bench.build_graph(torch, scale, "cuda:0", True)
    md = int((rp[1:] - rp[:-1]).max())
    for gt in (256, 512, 256, 512):
        capi.set_option("tc.gt2", gt)
        g = capi.DeviceGraph.adopt(rp, ci, md); g.prepare("tc")
        for _ in range(3): c = g.tc()
        ts = sorted(g.tc() and g.last_stats()[0] for _ in range(9))
        print(f"scale {scale} gt2={gt}: median {ts[4]:.3f} ms best {ts[0]:.3f} count {c}", flush=True)
        g.close()
    del rp, ci
