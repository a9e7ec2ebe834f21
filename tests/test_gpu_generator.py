"""The on-device graph generator (gm_gen_graph_*, csrc/gen.cu) produces exactly the graphs of the torch mirror
graphminer_b200/rmat.py (which the CPU tests, the oracle fixtures and the small bench sizes use), and the
device-side shard bounds agree with gm_host_shard_bounds."""
import numpy as np
import pytest
import torch

from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, shaped_graph

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nv,ns,seed,probs", [
    (1 << 10, 16 << 10, 0x5EED000A, (0.57, 0.19, 0.19, 0.05)),            # = rmat_graph(10)
    (1 << 14, 16 << 14, 0x5EED000E, (0.57, 0.19, 0.19, 0.05)),            # = rmat_graph(14)
    (3000, 40000, 0x5EED004C, (0.57, 0.19, 0.19, 0.05)),                  # LiveJournal shape, rejection of ids >= nv
    (4_847_571 // 64, 68_993_773 // 64, 0x5EED004C, (0.57, 0.19, 0.19, 0.05)),
    (65_608_366 // 512, 1_806_067_135 // 512, 0x5EED00F5, (0.45, 0.22, 0.22, 0.11)),   # Friendster shape
    (5, 40, 7, (0.25, 0.25, 0.25, 0.25)), (1, 10, 1, (0.57, 0.19, 0.19, 0.05)), (2, 0, 3, (0.57, 0.19, 0.19, 0.05)),
])
def test_generator_matches_torch_mirror(nv, ns, seed, probs):
    rp, ci = capi.generate_graph(nv, ns, seed, probs)
    want_rp, want_ci = shaped_graph(nv, ns, seed, probs=probs)
    assert rp.shape == (nv + 1,) and rp.dtype == torch.int64 and ci.dtype == torch.int32
    assert torch.equal(rp.cpu(), want_rp) and torch.equal(ci.cpu(), want_ci)


def test_generator_equals_rmat_graph():
    for scale in (8, 12):
        rp, ci = capi.generate_graph(1 << scale, 16 << scale, 0x5EED0000 + scale)
        want_rp, want_ci = rmat_graph(scale)
        assert torch.equal(rp.cpu(), want_rp) and torch.equal(ci.cpu(), want_ci)


def test_device_shard_bounds_match_host():
    rp, ci = rmat_graph(13)
    n_rp, n_ci = rp.numpy(), ci.numpy()
    with capi.DeviceGraph(n_rp, n_ci, 0) as g:
        for n in (1, 2, 3, 8):
            got = g.shard_bounds(n)
            want = [int(x) for x in capi.host_shard_bounds(n_rp, n_ci, n, balance=True)]
            assert got[0] == 0 and got[-1] == len(n_rp) - 1 and all(a <= b for a, b in zip(got, got[1:]))
            # double-precision prefix sums are accumulated in a different order on the device
            assert all(abs(a - b) <= 1 for a, b in zip(got, want)), (n, got, want)
