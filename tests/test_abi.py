"""CPU-side checks of the C ABI: the library loads, exports every declared symbol, the host-side
graph preparation matches the oracle, and compute entry points fail loudly without a GPU."""
import os
import re

import numpy as np
import pytest

import oracle
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gminer_b200.h")).read()
    declared = set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gm_graph_t"}
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    L = capi.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.gm_version() >= 100


def test_host_orient_edgelist_match_oracle(citeseer):
    rp, ci, _ = citeseer
    a = capi.host_orient(rp, ci)
    b = oracle.orient(rp, ci)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    for sb in (False, True):
        s1, d1 = capi.host_edgelist(rp, ci, sb)
        s2, d2 = oracle.edgelist(rp, ci, sb)
        assert np.array_equal(s1, s2) and np.array_equal(d1, d2)


def test_host_partition_matches_oracle_and_is_closed():
    rp, ci = (t.numpy() for t in rmat_graph(10))
    orp, oci, _ = oracle.orient(rp, ci)
    nv = len(orp) - 1
    total = 0
    bounds = capi.host_shard_bounds(orp, oci, 3, balance=False)
    assert list(bounds) == [0, 342, 684, 1024]                  # ceil(nv/3) chunks, graph_partition.cc:84-86
    for b, e in zip(bounds[:-1], bounds[1:]):
        got = capi.host_partition_part(orp, oci, int(b), int(e))
        exp = oracle.partition_part(orp, oci, int(b), int(e))
        for x, y in zip(got, exp):
            assert np.array_equal(x, y)
        srp, sci, idx, lb, le = got
        # the shard is closed for TC: counting its local source range reproduces the global range
        total += oracle.tc(srp, sci, (lb, le))
        assert oracle.tc(srp, sci, (lb, le)) == oracle.tc(orp, oci, (int(b), int(e)))
    assert total == oracle.tc(orp, oci)
    bal = capi.host_shard_bounds(orp, oci, 4, balance=True)
    assert bal[0] == 0 and bal[-1] == nv and np.all(np.diff(bal) >= 0)


def test_graph_file_round_trip(tmp_path, citeseer):
    rp, ci, md = citeseer
    prefix = str(tmp_path / "graph")
    capi.write_graph(prefix, rp, ci, md)
    rp2, ci2, md2 = capi.read_graph(prefix)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2) and md == md2
    with pytest.raises(capi.GMError):
        capi.read_graph(str(tmp_path / "missing"))


def test_compute_fails_loudly_without_gpu(citeseer):
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    rp, ci, md = citeseer
    with pytest.raises(capi.GMError):
        capi.DeviceGraph(rp, ci, md)
    with pytest.raises(capi.GMError):
        capi.tc_host(rp, ci, md)


def test_bad_arguments_are_errors_not_exits():
    with pytest.raises(capi.GMError):
        capi.set_option("no.such.option", "1")
    with pytest.raises(capi.GMError):
        capi.set_option("tc.algo", "bogus")
    # every value include/gminer_b200.h documents is accepted, numeric options are range-checked
    for v in ("auto", "rank", "hash", "hash_rev", "bs", "merge", "auto"):
        capi.set_option("tc.algo", v)
    for key, bad in (("sched.chunk", "-3"), ("sched.chunk", "x"), ("c4.small_max", "-2"), ("c4.cta_max", "1e3"),
                     ("tc.gt2", "300"), ("sup.gt2", "0"), ("clique.gt1", "1024"), ("batch.ring", "7")):
        with pytest.raises(capi.GMError):
            capi.set_option(key, bad)
    capi.set_option("sched.chunk", "0"); capi.set_option("c4.small_max", "-1")
    with pytest.raises(capi.GMError):
        capi.kclique_host(np.zeros(2, np.int64), np.zeros(0, np.int32), 9)
    with pytest.raises(capi.GMError):
        capi.sgl_host(np.zeros(2, np.int64), np.zeros(0, np.int32), "no-such-pattern")


def test_loader_sort_and_check(tmp_path):
    """Loader-side pieces: parallel positional read round-trips the reference format;
    gm_host_sort_neighbors = Graph::sort_neighbors; gm_host_check_sorted states the solvers' assumption."""
    import numpy as np
    rng = np.random.default_rng(3)
    nv = 5000
    deg = rng.integers(0, 40, nv)
    rp = np.zeros(nv + 1, np.int64); rp[1:] = np.cumsum(deg)
    rows = [np.sort(rng.choice(np.delete(np.arange(nv), v), d, replace=False)).astype(np.int32) for v, d in enumerate(deg)]
    ci = np.concatenate(rows) if len(rows) else np.zeros(0, np.int32)
    assert capi.check_sorted(rp, ci)
    shuffled = np.concatenate([rng.permutation(r) for r in rows])
    assert not capi.check_sorted(rp, shuffled)
    assert np.array_equal(capi.sort_neighbors(rp, shuffled), ci)
    loop = ci.copy(); loop[rp[7]] = 7 if deg[7] else loop[rp[7]]
    if deg[7]:
        assert not capi.check_sorted(rp, loop)
    prefix = str(tmp_path / "graph")
    capi.write_graph(prefix, rp, ci)
    rp2, ci2, md = capi.read_graph(prefix)
    assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and md == int(deg.max())
    with open(prefix + ".edge.bin", "r+b") as f:            # truncated file: loud failure, not a short read
        f.truncate(max(0, ci.nbytes - 4))
    with pytest.raises(capi.GMError):
        capi.read_graph(prefix)


def test_pinned_loader_roundtrip(tmp_path):
    """gm_host_alloc / gm_host_free + the loader reading straight into that memory (malloc fallback without a GPU)"""
    rng = np.random.default_rng(9)
    nv = 3000
    deg = rng.integers(0, 9, nv)
    rp = np.zeros(nv + 1, np.int64); np.cumsum(deg, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(nv, d, replace=False)) for d in deg]).astype(np.int32)
    prefix = str(tmp_path / "graph")
    capi.write_graph(prefix, rp, ci, int(deg.max()))
    got_rp, got_ci, md, pinned = capi.read_graph_pinned(prefix)
    assert np.array_equal(got_rp, rp) and np.array_equal(got_ci, ci) and md == int(deg.max())
    assert pinned == (capi.device_count() > 0)
    del got_rp, got_ci                                   # finalisers hand the arrays back to gm_host_free


def test_mapped_loader_roundtrip(tmp_path):
    """gm_host_map_graph: the reference's map_file path (custom_alloc.h:46-58) -- read-only mappings of the two
    arrays, registered with the CUDA driver when a device is present; truncated files are an error, not a fault"""
    rng = np.random.default_rng(10)
    nv = 2500
    deg = rng.integers(0, 7, nv)
    rp = np.zeros(nv + 1, np.int64); np.cumsum(deg, out=rp[1:])
    ci = np.concatenate([np.sort(rng.choice(nv, d, replace=False)) for d in deg]).astype(np.int32)
    prefix = str(tmp_path / "graph")
    capi.write_graph(prefix, rp, ci, int(deg.max()))
    got_rp, got_ci, md, pinned = capi.map_graph(prefix)
    assert np.array_equal(got_rp, rp) and np.array_equal(got_ci, ci) and md == int(deg.max())
    assert not got_rp.flags.writeable
    assert isinstance(pinned, bool) and (capi.device_count() > 0 or not pinned)
    orp, oci, omd = capi.host_orient(np.array(got_rp), np.array(got_ci))      # host logic reads the mapping directly
    assert len(orp) == nv + 1
    del got_rp, got_ci                                   # finalisers unmap
    with open(prefix + ".edge.bin", "r+b") as f:
        f.truncate(max(0, ci.nbytes - 8))
    with pytest.raises(capi.GMError):
        capi.map_graph(prefix)
