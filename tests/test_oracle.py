"""Pins the CPU oracle (oracle/gm_oracle.c) against the reference's golden vectors:
the README known-answer tables on the bundled citeseer / mico graphs."""
import json
import os

import numpy as np
import pytest

import oracle

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))


def test_citeseer_all(citeseer):
    rp, ci, md = citeseer
    k = KAT["citeseer"]
    assert md == 99 and len(rp) == 3313 and len(ci) == 9072
    orp, oci, omd = oracle.orient(rp, ci)
    assert oracle.tc(orp, oci) == k["tc"]
    assert oracle.kclique(orp, oci, 3) == k["tc"]
    assert oracle.kclique(orp, oci, 4) == k["clique4"]
    assert oracle.kclique(orp, oci, 5) == k["clique5"]
    for p in ("diamond", "rectangle", "house", "pentagon"):
        assert oracle.sgl(rp, ci, p) == k[p], p
    assert oracle.motif(rp, ci, 3) == k["motif3"]
    assert oracle.motif(rp, ci, 4) == k["motif4"]
    assert oracle.motif_formula(rp, ci, 3) == k["motif3"]
    assert oracle.motif_formula(rp, ci, 4) == k["motif4"]


def test_mico_fast(mico):
    rp, ci, md = mico
    k = KAT["mico"]
    assert md == 1359
    orp, oci, omd = oracle.orient(rp, ci)
    assert omd == 219 and len(oci) == 1080156
    assert oracle.tc(orp, oci) == k["tc"]
    assert oracle.kclique(orp, oci, 4) == k["clique4"]
    assert oracle.sgl(rp, ci, "diamond") == k["diamond"]
    assert oracle.motif(rp, ci, 3) == k["motif3"]
    assert oracle.motif_formula(rp, ci, 4) == k["motif4"]


@pytest.mark.slow
def test_mico_slow(mico):
    rp, ci, _ = mico
    k = KAT["mico"]
    assert oracle.sgl(rp, ci, "rectangle") == k["rectangle"]
    if os.environ.get("GM_SLOW_TESTS") == "1":            # ~30 s; the GPU suite checks this KAT too
        orp, oci, _ = oracle.orient(rp, ci)
        assert oracle.kclique(orp, oci, 5) == k["clique5"]


def test_range_split_is_additive(citeseer):
    rp, ci, _ = citeseer
    nv = len(rp) - 1
    orp, oci, _ = oracle.orient(rp, ci)
    cuts = [0, 500, 1700, nv]
    assert sum(oracle.tc(orp, oci, (a, b)) for a, b in zip(cuts, cuts[1:])) == 1166
    assert sum(oracle.sgl(rp, ci, "diamond", (a, b)) for a, b in zip(cuts, cuts[1:])) == 3730
    parts = [oracle.motif(rp, ci, 4, (a, b)) for a, b in zip(cuts, cuts[1:])]
    assert [sum(c) for c in zip(*parts)] == KAT["citeseer"]["motif4"]


def test_edgelist_semantics(citeseer):
    rp, ci, _ = citeseer
    src, dst = oracle.edgelist(rp, ci, sym_break=True)
    assert len(src) == len(ci) // 2 and np.all(src > dst)
    src, dst = oracle.edgelist(rp, ci, sym_break=False)
    assert len(src) == len(ci) and np.array_equal(dst, ci)


def test_reference_formula_shim_shards_add_up():
    """oracle/_ref/libgm_ref.so runs the reference's own formula 4-motif loop nest over a source range
    (bench.py's CPU baseline for the motif workload): raw sums of the shards + fix-up = the golden counts."""
    import json
    import os
    import numpy as np
    import oracle
    from graphminer_b200.rmat import rmat_graph
    if not oracle.have_ref() or not os.path.exists(os.path.join(oracle.REF_DIR, "libgm_ref.so")):
        import pytest
        pytest.skip("reference objects not built")
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rmat_counts.json")))["rmat12"]["motif4_formula"]
    rp, ci = (t.numpy() for t in rmat_graph(12))
    L = oracle.ref_lib()
    nv = len(rp) - 1
    h = L.gmr_graph_create(nv, rp, ci, int(np.diff(rp).max()))
    tot, raw = np.zeros(6, np.uint64), np.zeros(6, np.uint64)
    for b, e in ((0, 777), (777, 778), (778, nv)):
        L.gmr_motif4_formula_raw_range(h, b, e, raw)
        tot += raw
    L.gmr_graph_free(h)
    t = [int(x) for x in tot]
    t[4] = t[4] // 2 - t[5] * 6; t[2] = t[2] // 2 - t[4] * 2; t[1] = t[1] - t[3] * 4; t[0] = t[0] // 6 - t[2] // 3
    assert t == gold
