"""N>1 host logic on CPU: world_size-2 gloo processes shard by source range and all-reduce the
counts.  The per-shard counting is done by the oracle here (no GPU in this container); the GPU
version of the same flow is exercised by tests/test_gpu_parity.py (shards add up) and bench.py --gpus N."""
import json
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def _worker(rank, world, port, q):
    import oracle
    from graphminer_b200 import capi
    from graphminer_b200.dist import sharded_count, shard_bounds
    from tests.fixtures import load_fixture
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rp, ci, _ = load_fixture("citeseer")
        orp, oci, _ = capi.host_orient(rp, ci)
        res = {}
        res["tc"] = sharded_count(lambda b, e: oracle.tc(orp, oci, (b, e)), orp, oci)[0]
        res["clique4"] = sharded_count(lambda b, e: oracle.kclique(orp, oci, 4, (b, e)), orp, oci)[0]
        res["diamond"] = sharded_count(lambda b, e: oracle.sgl(rp, ci, "diamond", (b, e)), rp, ci)[0]
        res["motif4"] = sharded_count(lambda b, e: oracle.motif(rp, ci, 4, (b, e)), rp, ci, ncounts=6)
        # unbalanced (reference) split gives the same totals
        res["tc_eq"] = sharded_count(lambda b, e: oracle.tc(orp, oci, (b, e)), orp, oci, balance=False)[0]
        res["bounds"] = shard_bounds(orp, oci, world)
        # counts near 2^63 survive the int64 transport
        from graphminer_b200.dist import allreduce_counts
        res["big"] = allreduce_counts([2**62 + rank])[0]
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_reduce_to_kat():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k = KAT["citeseer"]
    for r in (0, 1):
        assert out[r]["tc"] == k["tc"] == out[r]["tc_eq"]
        assert out[r]["clique4"] == k["clique4"]
        assert out[r]["diamond"] == k["diamond"]
        assert out[r]["motif4"] == k["motif4"]
        assert out[r]["big"] == 2**63 + 1
        assert out[r]["bounds"][0] == 0 and out[r]["bounds"][-1] == 3312
