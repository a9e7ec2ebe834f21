import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, orient_dag
scale = int(sys.argv[1]) if len(sys.argv) > 1 else 22
what = sys.argv[2] if len(sys.argv) > 2 else "tc"
if what == "tc":
    rp, ci = rmat_graph(scale, device="cuda:0"); rp, ci = orient_dag(rp, ci)
else:
    from graphminer_b200.rmat import shaped_graph
    rp, ci = shaped_graph(4_847_571, 68_993_773, 0x5EED004C, device="cuda:0")
md = int((rp[1:] - rp[:-1]).max())
h_rp = torch.empty(rp.shape, dtype=rp.dtype, pin_memory=True); h_rp.copy_(rp)
h_ci = torch.empty(ci.shape, dtype=ci.dtype, pin_memory=True); h_ci.copy_(ci)
torch.cuda.synchronize()
n_rp, n_ci = h_rp.numpy(), h_ci.numpy()
for i in range(int(os.environ.get("GM_E2E_CALLS", "3"))):
    t0 = time.perf_counter()
    c = capi.tc_host(n_rp, n_ci, md) if what == "tc" else capi.sgl_host(n_rp, n_ci, "diamond", md)
    dt = time.perf_counter() - t0
    print(f"{what} host call {i}: {dt*1e3:.2f} ms count={c}", file=sys.stderr, flush=True)
