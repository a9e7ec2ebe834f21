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
