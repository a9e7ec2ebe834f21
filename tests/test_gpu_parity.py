"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, the reference README
known-answer tables (citeseer, mico) and the golden vectors of the reference's own OpenMP binaries
on generated R-MAT graphs.  Bit-exact: every quantity here is an integer count."""
import json
import os

import numpy as np
import pytest

import oracle
from graphminer_b200 import capi
from graphminer_b200.rmat import rmat_graph, shaped_graph

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(__file__)
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
GOLD = json.load(open(os.path.join(HERE, "golden", "rmat_counts.json")))
TC_ALGOS = ["auto", "rank", "hash", "hash_rev", "bs", "merge"]


@pytest.fixture(autouse=True)
def _reset_options():
    yield
    capi.set_option("tc.algo", "auto")
    capi.set_option("clique.algo", "auto")
    capi.set_option("sched.chunk", 0)
    capi.set_option("tc.shard", "source")
    capi.set_option("sgl.algo", "auto")
    capi.set_option("motif.algo", "auto")
    capi.set_option("c4.small_max", -1)
    capi.set_option("c4.mid_max", -1)
    capi.set_option("c4.cta_max", -1)
    capi.set_option("c4.hash", -1)
    capi.set_option("c4.persist", 1)
    capi.set_option("tc.short", 0)
    capi.set_option("tc.pipe", 0)
    capi.set_option("tc.flat", 5)
    capi.set_option("tc.hub", 65536)
    capi.set_option("tc.c1split", -1)
    capi.set_option("tc.c2split", 1)


def _graph(name):
    if name.startswith("rmat"):
        rp, ci = rmat_graph(int(name[4:]))
    else:
        rp, ci = shaped_graph(3000, 40000, 0x5EED004C)
    return rp.numpy(), ci.numpy()


def _dag(rp, ci):
    return capi.host_orient(rp, ci)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", TC_ALGOS)
@pytest.mark.parametrize("name", ["citeseer", "mico"])
def test_tc_kat(name, algo, request):
    rp, ci, _ = request.getfixturevalue(name)
    orp, oci, md = _dag(rp, ci)
    capi.set_option("tc.algo", algo)
    with capi.DeviceGraph(orp, oci, md) as g:
        assert g.tc() == KAT[name]["tc"]
        ms, launches = g.last_stats()
        assert launches >= 1 and ms > 0
        assert g.tc() == KAT[name]["tc"]          # cached aux structures, second call


def test_tc_alg_bytes_citeseer_mico(citeseer, mico):
    # SURVEY.md section 8(d) sanity values (with the rowptr term)
    for (rp, ci, _), want in ((citeseer, 127660), (mico, 280387976)):
        orp, oci, md = _dag(rp, ci)
        with capi.DeviceGraph(orp, oci, md) as g:
            g.tc()
            assert g.last_alg_bytes() == want


@pytest.mark.parametrize("algo", TC_ALGOS)
@pytest.mark.parametrize("name", ["rmat8", "rmat10", "rmat12", "rmat14", "rmat16", "shaped3000"])
def test_tc_golden(name, algo):
    rp, ci = _graph(name)
    orp, oci, md = _dag(rp, ci)
    capi.set_option("tc.algo", algo)
    with capi.DeviceGraph(orp, oci, md) as g:
        assert g.tc() == GOLD[name]["tc"]


@pytest.mark.parametrize("chunk", [1, 3, 1000000])
def test_tc_chunking_is_invisible(chunk):
    rp, ci = _graph("rmat12")
    orp, oci, md = _dag(rp, ci)
    capi.set_option("sched.chunk", chunk)
    for algo in ("rank", "hash", "hash_rev"):
        capi.set_option("tc.algo", algo)
        with capi.DeviceGraph(orp, oci, md) as g:
            assert g.tc() == GOLD["rmat12"]["tc"]


def test_tc_edge_cases():
    # empty graph, isolated vertices, a single triangle, a clique, a star (hub row much longer than the rest)
    def csr(n, edges):
        adj = [[] for _ in range(n)]
        for u, v in edges:
            adj[u].append(v); adj[v].append(u)
        rp = np.zeros(n + 1, np.int64)
        for i in range(n):
            rp[i + 1] = rp[i] + len(adj[i])
        ci = np.array([x for a in adj for x in sorted(a)], np.int32)
        return rp, ci
    cases = {
        "empty": (csr(5, []), 0),
        "triangle": (csr(4, [(0, 1), (1, 2), (0, 2)]), 1),
        "k6": (csr(6, [(i, j) for i in range(6) for j in range(i)]), 20),
        "star": (csr(300, [(0, i) for i in range(1, 300)]), 0),
        "wheel": (csr(200, [(0, i) for i in range(1, 200)] + [(i, i + 1) for i in range(1, 199)]), 198),
    }
    for name, ((rp, ci), want) in cases.items():
        orp, oci, md = _dag(rp, ci)
        assert oracle.tc(orp, oci) == want, name
        for algo in TC_ALGOS:
            capi.set_option("tc.algo", algo)
            with capi.DeviceGraph(orp, oci, max(md, 1)) as g:
                assert g.tc() == want, (name, algo)


def test_tc_big_rows_fall_back_correctly():
    # a row longer than the largest shared-memory table (d > 8192) exercises the global-search path
    n = 9500
    edges = [(0, i) for i in range(1, n)] + [(i, i + 1) for i in range(1, n - 1)]
    rp = np.zeros(n + 1, np.int64)
    adj = [[] for _ in range(n)]
    for u, v in edges:
        adj[u].append(v); adj[v].append(u)
    for i in range(n):
        rp[i + 1] = rp[i] + len(adj[i])
    ci = np.array([x for a in adj for x in sorted(a)], np.int32)
    # NOT oriented: feed the symmetric graph as if it were a DAG so row 0 keeps its 9499 entries;
    # sum over directed entries (u,v) of |N(u) ∩ N(v)| is still well defined and the oracle computes it
    want = oracle.tc(rp, ci)
    for algo in TC_ALGOS:
        capi.set_option("tc.algo", algo)
        with capi.DeviceGraph(rp, ci, n - 1) as g:
            assert g.tc() == want, algo


def test_tc_source_range_shards_add_up():
    rp, ci = _graph("rmat14")
    orp, oci, md = _dag(rp, ci)
    nv = len(orp) - 1
    bounds = capi.host_shard_bounds(orp, oci, 4, balance=True)
    for algo in TC_ALGOS:
        capi.set_option("tc.algo", algo)
        with capi.DeviceGraph(orp, oci, md) as g:
            parts = []
            for b, e in zip(bounds[:-1], bounds[1:]):
                g.set_source_range(int(b), int(e))
                parts.append(g.tc())
                assert parts[-1] == oracle.tc(orp, oci, (int(b), int(e)))
            assert sum(parts) == GOLD["rmat14"]["tc"]
            g.set_source_range(0, nv)
            assert g.tc() == GOLD["rmat14"]["tc"]


def test_tc_on_induced_partition():
    """graph_partition.cc semantics: each shard = range + 1-hop halo, relabelled; local range counted."""
    rp, ci = _graph("rmat12")
    orp, oci, md = _dag(rp, ci)
    bounds = capi.host_shard_bounds(orp, oci, 3, balance=False)
    total = 0
    for b, e in zip(bounds[:-1], bounds[1:]):
        srp, sci, idx, lb, le = capi.host_partition_part(orp, oci, int(b), int(e))
        with capi.DeviceGraph(srp, sci, 0) as g:
            g.set_source_range(lb, le)
            total += g.tc()
    assert total == GOLD["rmat12"]["tc"]


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo", ["list", "auto"])
def test_kclique_kat(citeseer, mico, algo):
    capi.set_option("clique.algo", algo)
    rp, ci, _ = citeseer
    orp, oci, md = _dag(rp, ci)
    with capi.DeviceGraph(orp, oci, md) as g:
        assert g.kclique(3) == KAT["citeseer"]["tc"]
        assert g.kclique(4) == KAT["citeseer"]["clique4"]
        assert g.kclique(5) == KAT["citeseer"]["clique5"]
    rp, ci, _ = mico
    orp, oci, md = _dag(rp, ci)
    with capi.DeviceGraph(orp, oci, md) as g:
        assert g.kclique(4) == KAT["mico"]["clique4"]
        assert g.kclique(5) == KAT["mico"]["clique5"]


@pytest.mark.parametrize("algo", ["list", "auto"])
@pytest.mark.parametrize("name", ["rmat8", "rmat10", "rmat12", "rmat14", "shaped3000"])
def test_kclique_golden(name, algo):
    capi.set_option("clique.algo", algo)
    rp, ci = _graph(name)
    orp, oci, md = _dag(rp, ci)
    with capi.DeviceGraph(orp, oci, md) as g:
        assert g.kclique(4) == GOLD[name]["clique4"]
        assert g.kclique(5) == GOLD[name]["clique5"]


@pytest.mark.parametrize("algo", ["list", "auto"])
def test_kclique_deep_matches_closed_form(algo):
    # K_n contains C(n,k) k-cliques: checks k = 6, 7, 8 (the oracle stops at 5, automine_omp.h:159-183)
    from math import comb
    capi.set_option("clique.algo", algo)
    n = 14
    rp = np.arange(0, n * (n - 1) + 1, n - 1, dtype=np.int64)
    ci = np.array([j for i in range(n) for j in range(n) if j != i], np.int32)
    orp, oci, md = _dag(rp, ci)
    with capi.DeviceGraph(orp, oci, md) as g:
        for k in range(3, 9):
            assert g.kclique(k) == comb(n, k), k
        with pytest.raises(capi.GMError):
            g.kclique(9)


def test_kclique_shards_add_up():
    rp, ci = _graph("rmat12")
    orp, oci, md = _dag(rp, ci)
    bounds = capi.host_shard_bounds(orp, oci, 3, balance=True)
    for algo in ("list", "auto"):
        capi.set_option("clique.algo", algo)
        with capi.DeviceGraph(orp, oci, md) as g:
            tot = 0
            for b, e in zip(bounds[:-1], bounds[1:]):
                g.set_source_range(int(b), int(e))
                tot += g.kclique(4)
            assert tot == GOLD["rmat12"]["clique4"]


# ---------------------------------------------------------------------------------------------
def test_sgl_kat(citeseer, mico):
    rp, ci, md = citeseer
    with capi.DeviceGraph(rp, ci, md) as g:
        for p in ("diamond", "rectangle", "house", "pentagon"):
            assert g.sgl(p) == KAT["citeseer"][p], p
        with pytest.raises(capi.GMError):
            g.sgl("dumbbell")
    rp, ci, md = mico
    with capi.DeviceGraph(rp, ci, md) as g:
        assert g.sgl("diamond") == KAT["mico"]["diamond"]
        assert g.sgl("rectangle") == KAT["mico"]["rectangle"]


@pytest.mark.parametrize("name", ["rmat8", "rmat10", "rmat12", "shaped3000"])
def test_sgl_golden(name):
    rp, ci = _graph(name)
    with capi.DeviceGraph(rp, ci, 0) as g:
        for p in ("diamond", "rectangle", "house", "pentagon"):
            assert g.sgl(p) == GOLD[name][p], (name, p)


def test_sgl_shards_add_up():
    rp, ci = _graph("rmat10")
    nv = len(rp) - 1
    with capi.DeviceGraph(rp, ci, 0) as g:
        for p in ("diamond", "rectangle", "house", "pentagon"):
            tot = 0
            for b, e in ((0, 300), (300, 301), (301, nv)):
                g.set_source_range(b, e)
                got = g.sgl(p)
                assert got == oracle.sgl(rp, ci, p, (b, e)), (p, b, e)
                tot += got
            assert tot == GOLD["rmat10"][p]


@pytest.mark.parametrize("algo", ["support", "list"])
def test_diamond_both_algorithms(algo, citeseer, mico):
    """diamond by per-edge triangle supports on the device-oriented DAG (support.cu) and by the
    warp-per-edge operator-API kernel: KATs, golden R-MAT / shaped counts, per-shard oracle equality,
    degenerate graphs."""
    capi.set_option("sgl.algo", algo)
    for (rp, ci, md), name in ((citeseer, "citeseer"), (mico, "mico")):
        with capi.DeviceGraph(rp, ci, md) as g:
            assert g.sgl("diamond") == KAT[name]["diamond"]
            assert g.sgl("diamond") == KAT[name]["diamond"]       # cached structures, second call
    for name in ("rmat8", "rmat12", "rmat14", "shaped3000"):
        rp, ci = _graph(name)
        with capi.DeviceGraph(rp, ci, 0) as g:
            want = GOLD[name]["diamond"] if name in GOLD else oracle.sgl(rp, ci, "diamond")
            assert g.sgl("diamond") == want, name
    rp, ci = _graph("rmat10")
    nv = len(rp) - 1
    with capi.DeviceGraph(rp, ci, 0) as g:
        tot = 0
        for b, e in ((0, 300), (300, 301), (301, nv)):
            g.set_source_range(b, e)
            got = g.sgl("diamond")
            assert got == oracle.sgl(rp, ci, "diamond", (b, e)), (b, e)
            tot += got
        assert tot == GOLD["rmat10"]["diamond"]
    # K_n: every edge lies in n-2 triangles -> C(n,2) * C(n-2,2) diamonds; star / empty graph: none
    n = 40
    rp = np.arange(0, n * (n - 1) + 1, n - 1, dtype=np.int64)
    ci = np.concatenate([np.delete(np.arange(n, dtype=np.int32), i) for i in range(n)])
    with capi.DeviceGraph(rp, ci, 0) as g:
        assert g.sgl("diamond") == (n * (n - 1) // 2) * ((n - 2) * (n - 3) // 2)
    m = 50
    rp = np.concatenate([[0], np.arange(m, 2 * m + 1)]).astype(np.int64)
    ci = np.concatenate([np.arange(1, m + 1), np.zeros(m)]).astype(np.int32)
    with capi.DeviceGraph(rp, ci, 0) as g:
        assert g.sgl("diamond") == 0


# ---------------------------------------------------------------------------------------------
def test_motif_kat(citeseer, mico):
    rp, ci, md = citeseer
    with capi.DeviceGraph(rp, ci, md) as g:
        assert g.motif(3) == KAT["citeseer"]["motif3"]
        assert g.motif(4) == KAT["citeseer"]["motif4"]
        assert g.motif(3, formula=True) == KAT["citeseer"]["motif3"]
        assert g.motif(4, formula=True) == KAT["citeseer"]["motif4"]
        with pytest.raises(capi.GMError):
            g.motif(5)
    rp, ci, md = mico
    with capi.DeviceGraph(rp, ci, md) as g:
        assert g.motif(3) == KAT["mico"]["motif3"]
        assert g.motif(4, formula=True) == KAT["mico"]["motif4"]
        assert g.motif(4) == KAT["mico"]["motif4"]


@pytest.mark.parametrize("name", ["rmat8", "rmat10", "rmat12", "shaped3000"])
def test_motif_golden(name):
    rp, ci = _graph(name)
    with capi.DeviceGraph(rp, ci, 0) as g:
        assert g.motif(3) == GOLD[name]["motif3"]
        assert g.motif(4) == GOLD[name]["motif4"]
        assert g.motif(4, formula=True) == GOLD[name]["motif4_formula"]
        assert g.motif(3, formula=True) == GOLD[name]["motif3"]


def test_motif_shards_add_up():
    rp, ci = _graph("rmat10")
    nv = len(rp) - 1
    with capi.DeviceGraph(rp, ci, 0) as g:
        base = np.zeros(6, np.uint64); raw = np.zeros(6, np.uint64)
        for b, e in ((0, 400), (400, nv)):
            g.set_source_range(b, e)
            got = g.motif(4)
            assert got == oracle.motif(rp, ci, 4, (b, e))
            base += np.array(got, np.uint64)
            raw += np.array(g.motif(4, formula=True, raw=True), np.uint64)
        assert [int(x) for x in base] == GOLD["rmat10"]["motif4"]
        assert capi.motif_formula_finish(4, raw) == GOLD["rmat10"]["motif4"]


# ---------------------------------------------------------------------------------------------
def test_host_entry_points(citeseer):
    rp, ci, md = citeseer
    orp, oci, omd = _dag(rp, ci)
    k = KAT["citeseer"]
    assert capi.tc_host(orp, oci, omd) == k["tc"]
    assert capi.kclique_host(orp, oci, 4, omd) == k["clique4"]
    assert capi.sgl_host(rp, ci, "diamond", md) == k["diamond"]
    assert capi.motif_host(rp, ci, 4, False, md) == k["motif4"]
    assert capi.motif_host(rp, ci, 4, True, md) == k["motif4"]
    assert capi.motif_host(rp, ci, 3, True, md) == k["motif3"]
    # asking for more GPUs than present clamps (triangle/multigpu.cu:28-30) and still counts right
    assert capi.tc_host(orp, oci, omd, n_gpus=64) == k["tc"]
    assert capi.motif_host(rp, ci, 4, True, md, n_gpus=64) == k["motif4"]


def test_device_side_results(citeseer):
    """gm_graph_set_result_buffer: asynchronous solvers leave their counts on the device in stream order
    (what bench.py --gpus N chains its NCCL all-reduce on); NULL restores the synchronous path."""
    import torch
    rp, ci, _ = citeseer
    orp, oci, md = _dag(rp, ci)
    k = KAT["citeseer"]
    res = torch.full((8,), -1, dtype=torch.int64, device="cuda:0")
    with capi.DeviceGraph(orp, oci, md) as g:
        g.set_result_buffer(res)
        assert g.tc() == 0                         # host result untouched in asynchronous mode
        g.kclique(4)                               # second pass queued behind the first, no sync in between
        torch.cuda.synchronize()
        assert int(res[0]) == k["clique4"]
        g.tc(); torch.cuda.synchronize()
        assert int(res[0]) == k["tc"]
        ms, launches = g.last_stats()
        assert ms > 0 and launches >= 1
        g.set_result_buffer(None)
        assert g.tc() == k["tc"]
    with capi.DeviceGraph(rp, ci, 0) as g:
        g.set_result_buffer(res)
        g.motif(4, formula=True, raw=True); torch.cuda.synchronize()
        assert capi.motif_formula_finish(4, [int(x) for x in res[:6].tolist()]) == k["motif4"]
        with pytest.raises(capi.GMError):
            g.motif(4, formula=True)


def test_tc_destination_sharding_adds_up():
    """tc.shard=dest: a shard owns the edges whose DESTINATION lies in its range (each root's table is then
    built on one shard only); over a partition of the vertex set the shards still add up to the oracle."""
    rp, ci = _graph("rmat14")
    orp, oci, md = _dag(rp, ci)
    want = oracle.tc(orp, oci)
    nv = len(orp) - 1
    capi.set_option("tc.shard", "dest")
    cuts = [0, nv // 5, nv // 2, nv - 7, nv]
    total = 0
    with capi.DeviceGraph(orp, oci, md) as g:
        for b, e in zip(cuts[:-1], cuts[1:]):
            g.set_source_range(b, e)
            total += g.tc()
    assert total == want


@pytest.mark.parametrize("algo", ["fast", "list"])
def test_motif4_formula_both_algorithms(algo, citeseer, mico):
    """4-motif formula: supports + wedge-pair 4-cycles + bit-matrix 4-cliques on the DAG (cycle4.cu) and the
    warp-per-edge operator-API kernel; KATs, golden graphs, shards (raw sums add up before the fix-up),
    4-cycle-only graphs that exercise the chord correction."""
    capi.set_option("motif.algo", algo)
    for (rp, ci, md), name in ((citeseer, "citeseer"), (mico, "mico")):
        with capi.DeviceGraph(rp, ci, md) as g:
            assert g.motif(4, formula=True) == KAT[name]["motif4"]
            assert g.motif(4, formula=True) == KAT[name]["motif4"]
    for name in ("rmat8", "rmat10", "rmat12", "rmat14", "shaped3000"):
        rp, ci = _graph(name)
        want = GOLD[name].get("motif4_formula") or GOLD[name].get("motif4")
        with capi.DeviceGraph(rp, ci, 0) as g:
            assert g.motif(4, formula=True) == want, name
    rp, ci = _graph("rmat12")
    nv = len(rp) - 1
    with capi.DeviceGraph(rp, ci, 0) as g:
        raw = np.zeros(6, np.uint64)
        for b, e in ((0, 1000), (1000, 1001), (1001, nv)):
            g.set_source_range(b, e)
            raw += np.array(g.motif(4, formula=True, raw=True), np.uint64)      # wraps mod 2^64 by design
        assert capi.motif_formula_finish(4, raw) == GOLD["rmat12"]["motif4_formula"]
    # complete bipartite K_{a,b}: C(a,2)*C(b,2) chordless 4-cycles, no triangles
    a, b = 7, 9
    rp = np.concatenate([[0], np.cumsum([b] * a + [a] * b)]).astype(np.int64)
    ci = np.concatenate([np.arange(a, a + b)] * a + [np.arange(a)] * b).astype(np.int32)
    with capi.DeviceGraph(rp, ci, 0) as g:
        got = g.motif(4, formula=True)
        assert got[3] == (a * (a - 1) // 2) * (b * (b - 1) // 2) and got[4] == 0 and got[5] == 0
        assert got == oracle.motif(rp, ci, 4)
    # K_6: 15 four-cliques, no chordless cycle, no induced diamond
    n = 6
    rp = np.arange(0, n * (n - 1) + 1, n - 1, dtype=np.int64)
    ci = np.concatenate([np.delete(np.arange(n, dtype=np.int32), i) for i in range(n)])
    with capi.DeviceGraph(rp, ci, 0) as g:
        assert g.motif(4, formula=True) == [0, 0, 0, 0, 0, 15]


@pytest.mark.parametrize("hash_tier", [0, 1])
@pytest.mark.parametrize("small_max,cta_max,mid_max", [(0, -1, -1), (0, 0, -1), (0, 0, 0), (16, 100, 300), (512, 600, 2000)])
def test_motif4_cycle_tiers(small_max, cta_max, mid_max, hash_tier):
    """the four 4-cycle tiers (warp table / CTA table / cluster on a dense array or -- large graphs -- on per-root
    hash tables / whole-grid dense array) must agree: thresholds are pushed down so that small graphs reach every
    code path"""
    capi.set_option("motif.algo", "fast")
    capi.set_option("c4.hash", hash_tier)
    capi.set_option("c4.small_max", small_max)
    capi.set_option("c4.cta_max", cta_max)
    capi.set_option("c4.mid_max", mid_max)
    for name in ("rmat10", "rmat14", "shaped3000"):
        rp, ci = _graph(name)
        with capi.DeviceGraph(rp, ci, 0) as g:
            assert g.motif(4, formula=True) == GOLD[name]["motif4_formula"], name


def test_diamond_partitioned_support_exchange():
    """the multi-GPU diamond path on one device: N logical shards each enumerate the triangles of their root
    range (gm_sgl_support_begin), the support arrays are summed (what the NCCL all-reduce does), every shard
    sums C(t,2) over its own edges (gm_sgl_support_finish) -- per-shard oracle equality and the total."""
    import torch
    rp, ci = _graph("rmat12")
    nv = len(rp) - 1
    cuts = [0, nv // 3, nv // 3 + 1, nv - 100, nv]
    shards = []
    for b, e in zip(cuts[:-1], cuts[1:]):
        g = capi.DeviceGraph(rp, ci, 0)
        g.set_source_range(b, e)
        g.sgl_support_begin()
        shards.append((g, b, e))
    torch.cuda.synchronize()
    sups = [g.support_tensor() for g, _, _ in shards]
    total_sup = torch.stack(sups).sum(0).to(torch.int32)
    assert int(total_sup.sum()) == 3 * oracle.tc(*capi.host_orient(rp, ci)[:2])      # every triangle marks 3 edges
    got = 0
    for (g, b, e), sup in zip(shards, sups):
        sup.copy_(total_sup)
        c = g.sgl_support_finish()
        assert c == oracle.sgl(rp, ci, "diamond", (b, e)), (b, e)
        got += c
        # the handle still serves the single-GPU solvers afterwards (child switches back to the full pass)
        assert g.sgl("diamond") == oracle.sgl(rp, ci, "diamond", (b, e))
        g.close()
    assert got == GOLD["rmat12"]["diamond"]


@pytest.mark.parametrize("algo", ["auto", "list"])
def test_rectangle_both_algorithms(algo, citeseer, mico):
    """sgl rectangle (every 4-cycle once): wedge-pair counting on the DAG (full source range) and the
    warp-per-edge operator-API kernel; with a partial range both settings take the operator-API kernel."""
    capi.set_option("sgl.algo", algo)
    for (rp, ci, md), name in ((citeseer, "citeseer"), (mico, "mico")):
        with capi.DeviceGraph(rp, ci, md) as g:
            assert g.sgl("rectangle") == KAT[name]["rectangle"]
            assert g.sgl("diamond") == KAT[name]["diamond"]          # the two fast paths share the child handle
            assert g.sgl("rectangle") == KAT[name]["rectangle"]
    for name in ("rmat8", "rmat10", "rmat12", "shaped3000"):
        rp, ci = _graph(name)
        with capi.DeviceGraph(rp, ci, 0) as g:
            assert g.sgl("rectangle") == GOLD[name]["rectangle"], name
            nv = len(rp) - 1
            g.set_source_range(nv // 2, nv)
            assert g.sgl("rectangle") == oracle.sgl(rp, ci, "rectangle", (nv // 2, nv))
            g.set_source_range(0, nv)
            assert g.sgl("rectangle") == GOLD[name]["rectangle"]
    a, b = 6, 11                                                     # K_{a,b}: C(a,2) * C(b,2) four-cycles
    rp = np.concatenate([[0], np.cumsum([b] * a + [a] * b)]).astype(np.int64)
    ci = np.concatenate([np.arange(a, a + b)] * a + [np.arange(a)] * b).astype(np.int32)
    with capi.DeviceGraph(rp, ci, 0) as g:
        assert g.sgl("rectangle") == (a * (a - 1) // 2) * (b * (b - 1) // 2)
    n = 9                                                            # K_n: 3 * C(n,4) four-cycles
    rp = np.arange(0, n * (n - 1) + 1, n - 1, dtype=np.int64)
    ci = np.concatenate([np.delete(np.arange(n, dtype=np.int32), i) for i in range(n)])
    with capi.DeviceGraph(rp, ci, 0) as g:
        assert g.sgl("rectangle") == 3 * (n * (n - 1) * (n - 2) * (n - 3) // 24)


# ---- round 2 ------------------------------------------------------------------------------------
def _complete_dag(n):
    """K_n already oriented: all degrees are equal, so the (degree, id) order is the id order"""
    deg = np.arange(n - 1, -1, -1, dtype=np.int64)
    rp = np.zeros(n + 1, np.int64); np.cumsum(deg, out=rp[1:])
    ci = np.concatenate([np.arange(i + 1, n, dtype=np.int32) for i in range(n)]) if n > 1 else np.zeros(0, np.int32)
    return rp, ci, n - 1


def test_tc_ranked_rows_of_every_size_class():
    """rank.cu sorts the relabelled rows per size class: registers (d <= 32), a warp in shared memory (<= 256),
    a CTA (<= 4096) and the segmented radix sort beyond; K_4500's DAG has rows of every length 0..4499."""
    n = 4500
    rp, ci, md = _complete_dag(n)
    want = n * (n - 1) * (n - 2) // 6
    for algo in ("rank", "merge", "hash_rev"):
        capi.set_option("tc.algo", algo)
        with capi.DeviceGraph(rp, ci, md) as g:
            assert g.tc() == want, algo
    # a scrambled labelling of the same DAG: ranks differ from ids, rows must really be sorted
    rng = np.random.default_rng(5)
    m = 1500
    perm = rng.permutation(m).astype(np.int64)
    rows = [[] for _ in range(m)]
    for i in range(m):
        for j in range(i + 1, m):
            a, b = int(perm[i]), int(perm[j])
            rows[min(a, b)].append(max(a, b))            # equal degrees: orientation by id
    rp2 = np.zeros(m + 1, np.int64)
    for i in range(m):
        rp2[i + 1] = rp2[i] + len(rows[i])
    ci2 = np.array([x for r in rows for x in sorted(r)], np.int32)
    capi.set_option("tc.algo", "rank")
    with capi.DeviceGraph(rp2, ci2, m - 1) as g:
        assert g.tc() == m * (m - 1) * (m - 2) // 6


def test_max_degree_is_recomputed_when_the_caller_value_is_stale(mico):
    rp, ci, md = mico
    orp, oci, omd = _dag(rp, ci)
    capi.set_option("clique.algo", "list")                # per-warp frontiers are sized from max_degree
    with capi.DeviceGraph(orp, oci, 1) as g:              # stale meta.txt value
        assert g.info()["max_degree"] == omd
        assert g.kclique(4) == KAT["mico"]["clique4"]


@pytest.mark.parametrize("n,k", [(150, 6), (100, 7), (72, 8)])
def test_kclique_dense_neighbourhood_beyond_32_bits(n, k):
    """C(n,k) > 2^32 on one dense neighbourhood: the per-root / per-lane accumulators of the bit-matrix
    kernel are 64-bit (round 1 held a root's whole sub-tree in uint32)"""
    import math
    rp, ci, md = _complete_dag(n)
    want = math.comb(n, k)
    assert want > 2 ** 32
    with capi.DeviceGraph(rp, ci, md) as g:
        assert g.kclique(k) == want


def _need_devices(n):
    if capi.device_count() < n:
        pytest.skip(f"needs {n} CUDA devices")


def test_host_entry_points_on_two_devices(citeseer, mico):
    """gm_*_host(n_gpus=2): one host thread per device, sharded upload + NCCL all-gather, support exchange for
    diamond and the formula 4-motif, NCCL all-reduce of the counts (solvers.cu run_shard)"""
    _need_devices(2)
    for name, (rp, ci, md) in (("citeseer", citeseer), ("mico", mico)):
        k = KAT[name]
        orp, oci, omd = _dag(rp, ci)
        for rep in range(2):                              # second call: cached communicators and handle resources
            assert capi.tc_host(orp, oci, omd, n_gpus=2) == k["tc"]
            assert capi.kclique_host(orp, oci, 4, omd, n_gpus=2) == k["clique4"]
            assert capi.sgl_host(rp, ci, "diamond", md, n_gpus=2) == k["diamond"]
            assert capi.motif_host(rp, ci, 4, True, md, n_gpus=2) == k["motif4"]
            assert capi.motif_host(rp, ci, 3, False, md, n_gpus=2) == k["motif3"]
    rp, ci, md = citeseer
    assert capi.sgl_host(rp, ci, "rectangle", md, n_gpus=2) == KAT["citeseer"]["rectangle"]
    for hub in (1024, 65536):                             # hybrid rows on both sides of the hub split, sharded
        capi.set_option("tc.hub", hub)
        grp, gci = _graph("rmat16")
        orp, oci, omd = _dag(grp, gci)
        assert capi.tc_host(orp, oci, omd, n_gpus=2) == GOLD["rmat16"]["tc"]
    capi.set_option("sgl.algo", "list"); capi.set_option("motif.algo", "list")
    assert capi.sgl_host(rp, ci, "diamond", md, n_gpus=2) == KAT["citeseer"]["diamond"]
    assert capi.motif_host(rp, ci, 4, True, md, n_gpus=2) == KAT["citeseer"]["motif4"]


def test_second_device_runs_every_kernel_family(citeseer):
    """per-device shared-memory opt-in attributes (a process-wide "already set" flag broke device != 0)"""
    _need_devices(2)
    import torch
    rp, ci, md = citeseer
    orp, oci, omd = _dag(rp, ci)
    with capi.DeviceGraph(orp, oci, omd, device=0) as g0, capi.DeviceGraph(orp, oci, omd, device=1) as g1:
        assert g0.tc() == g1.tc() == KAT["citeseer"]["tc"]
        assert g1.kclique(5) == KAT["citeseer"]["clique5"]
    with capi.DeviceGraph(rp, ci, md, device=1) as g1:
        assert g1.motif(4, formula=True) == KAT["citeseer"]["motif4"]
        assert g1.sgl("diamond") == KAT["citeseer"]["diamond"]
    a = np.unique(np.random.default_rng(0).integers(0, 4000, 900)).astype(np.int32)
    b = np.unique(np.random.default_rng(1).integers(0, 4000, 700)).astype(np.int32)
    for dev in (0, 1):
        with torch.cuda.device(dev):
            pool = torch.from_numpy(np.concatenate([a, b, np.zeros(8, np.int32)])).cuda()
            ao = torch.tensor([0], dtype=torch.int64, device="cuda"); bo = torch.tensor([len(a)], dtype=torch.int64, device="cuda")
            al = torch.tensor([len(a)], dtype=torch.int32, device="cuda"); bl = torch.tensor([len(b)], dtype=torch.int32, device="cuda")
            for algo in ("auto", "merge", "hash", "gallop", "bsearch"):
                got = int(capi.intersect_batch(pool, ao, al, bo, bl, algo=algo).cpu()[0])
                assert got == oracle.intersection_num(a, b), (dev, algo)


def test_motif4_partitioned_support_exchange():
    """gm_motif_support_begin / finish: shards enumerate the triangles of their root range only; summing the
    support arrays (what the NCCL all-reduce does) and finishing every shard gives the whole-graph counts"""
    import torch
    rp, ci = _graph("rmat12")
    nv = len(rp) - 1
    want = oracle.motif_formula(rp, ci, 4)
    bounds = [0, nv // 3, nv // 2, nv]
    gs = [capi.DeviceGraph(rp, ci, 0) for _ in range(3)]
    try:
        sups = []
        for g, b, e in zip(gs, bounds[:-1], bounds[1:]):
            g.set_source_range(b, e)
            g.motif_support_begin()
            sups.append(g.support_tensor())
        torch.cuda.synchronize()
        total = sum(s.clone() for s in sups)
        raw = [0] * 6
        for g, s in zip(gs, sups):
            s.copy_(total)
            raw = [(x + y) & (2 ** 64 - 1) for x, y in zip(raw, g.motif_support_finish())]
        assert capi.motif_formula_finish(4, raw) == want
        # and again on the same handles (cached structures)
        for g in gs:
            g.motif_support_begin()
        torch.cuda.synchronize()
        total = sum(g.support_tensor().clone() for g in gs)
        raw = [0] * 6
        for g in gs:
            g.support_tensor().copy_(total)
            raw = [(x + y) & (2 ** 64 - 1) for x, y in zip(raw, g.motif_support_finish())]
        assert capi.motif_formula_finish(4, raw) == want
    finally:
        for g in gs:
            g.close()


@pytest.mark.parametrize("short_max", [1, 8, 32, 1000])
@pytest.mark.parametrize("pipe", [0, 1])
def test_tc_stream_loop_variants(short_max, pipe):
    """tc.short (suffixes walked by one lane each) and tc.pipe (cross-partner prefetch) are schedules of the same
    count: every golden graph, the ranked and the unranked table kernels, and the support / clique kernels that
    share the partner records stay exact"""
    capi.set_option("tc.short", short_max)
    capi.set_option("tc.pipe", pipe)
    for name in ("rmat8", "rmat12", "rmat14", "shaped3000"):
        rp, ci = _graph(name)
        orp, oci, md = _dag(rp, ci)
        for algo in ("rank", "hash", "hash_rev"):
            capi.set_option("tc.algo", algo)
            with capi.DeviceGraph(orp, oci, md) as g:
                assert g.tc() == GOLD[name]["tc"], (name, algo)
    rp, ci, _ = _complete_dag(300)
    capi.set_option("tc.algo", "rank")
    with capi.DeviceGraph(rp, ci, 299) as g:
        assert g.tc() == 300 * 299 * 298 // 6


def test_c4_persisting_window_is_only_a_hint():
    for persist in (0, 1):
        capi.set_option("c4.persist", persist)
        capi.set_option("c4.small_max", 0); capi.set_option("c4.cta_max", 0)
        rp, ci = _graph("rmat12")
        with capi.DeviceGraph(rp, ci, 0) as g:
            assert g.motif(4, formula=True) == GOLD["rmat12"]["motif4_formula"]


@pytest.mark.parametrize("hash_tier", [0, 1])
@pytest.mark.parametrize("small_max,cta_max,mid_max", [(-1, -1, -1), (0, -1, -1), (0, 0, -1), (0, 0, 0), (16, 100, 300)])
def test_house_on_the_dag_machinery(small_max, cta_max, mid_max, hash_tier, citeseer, mico):
    """house = sum_e [t(e) sq(e) - 2 t(e)^2 + 2 t(e)]: supports + per-edge 4-cycle counts attributed by every tier
    of the wedge-pair kernel (cycle4.cu ATTR) against the README counts, the reference OMP goldens and the
    operator-API kernel"""
    capi.set_option("c4.small_max", small_max); capi.set_option("c4.cta_max", cta_max); capi.set_option("c4.mid_max", mid_max)
    capi.set_option("c4.hash", hash_tier)
    cases = [(citeseer[0], citeseer[1], KAT["citeseer"]["house"])]
    if small_max == -1:
        cases.append((mico[0], mico[1], KAT["mico"]["house"]))
    for name in ("rmat8", "rmat10", "shaped3000"):
        rp, ci = _graph(name)
        cases.append((rp, ci, GOLD[name]["house"] if "house" in GOLD[name] else oracle.sgl(rp, ci, "house")))
    for rp, ci, want in cases:
        with capi.DeviceGraph(rp, ci, 0) as g:
            assert g.sgl("house") == want
            assert g.sgl("house") == want                      # cached structures, cleared tables
            assert g.sgl("rectangle") == oracle.sgl(rp, ci, "rectangle") if len(rp) < 5000 else True
    capi.set_option("sgl.algo", "list")
    rp, ci = _graph("rmat10")
    with capi.DeviceGraph(rp, ci, 0) as g:
        want = oracle.sgl(rp, ci, "house")
        assert g.sgl("house") == want
        g.set_source_range(10, 500)                              # a shard keeps the operator-API kernel
        capi.set_option("sgl.algo", "auto")
        assert g.sgl("house") == oracle.sgl(rp, ci, "house", (10, 500))


@pytest.mark.parametrize("hub", [16, 256, 4096, 65536])
def test_tc_hybrid_rows_split_at_every_hub_range(hub):
    """tc.flat=5 keeps the top `tc.hub` ranks of every row as 16-rank bitmap blocks and the rest as hashed keys;
    the golden graphs are smaller than the default range (everything a hub), so the range is shrunk to put
    roots and partners on both sides of the split.  The other stream loops (0: per record, 1: flat windows,
    4: scaled keys) count the same on the SAME handle: switching rebuilds the plain partner records."""
    capi.set_option("tc.hub", hub)
    for name in ("rmat8", "rmat12", "rmat14", "rmat16", "shaped3000"):
        rp, ci = _graph(name)
        orp, oci, md = _dag(rp, ci)
        capi.set_option("tc.flat", 5)
        capi.set_option("tc.algo", "rank")
        with capi.DeviceGraph(orp, oci, md) as g:
            assert g.tc() == GOLD[name]["tc"], (name, hub)
            for flat in (1, 4, 0, 5):
                capi.set_option("tc.flat", flat)
                assert g.tc() == GOLD[name]["tc"], (name, hub, flat)
            for c1, c2 in ((0, 0), (1, 1), (0, 1), (1, 0)):          # group widths of the hybrid kernel's size classes
                capi.set_option("tc.c1split", c1); capi.set_option("tc.c2split", c2)
                assert g.tc() == GOLD[name]["tc"], (name, hub, c1, c2)
            capi.set_option("tc.c1split", -1); capi.set_option("tc.c2split", 1)
            capi.set_option("tc.algo", "merge")
            assert g.tc() == GOLD[name]["tc"], (name, hub, "merge")
    capi.set_option("tc.flat", 5)
    capi.set_option("tc.algo", "rank")
    n = 700                                                  # K_700: rows of every length, all blocks dense
    rp, ci, md = _complete_dag(n)
    with capi.DeviceGraph(rp, ci, md) as g:
        for c1, c2 in ((0, 0), (1, 1)):                      # with a small hub range: key tables beyond both small configurations
            capi.set_option("tc.c1split", c1); capi.set_option("tc.c2split", c2)
            assert g.tc() == n * (n - 1) * (n - 2) // 6, (hub, c1, c2)


def test_tc_hybrid_rows_beyond_the_default_hub_range():
    """R-MAT scale 18 has 4x more vertices than the default hub range: keys and bitmap blocks both in play,
    checked against the operator-API kernel and the plain ranked kernels, whole graph and shards (by source and
    by destination)"""
    rp, ci = rmat_graph(18)
    orp, oci, md = capi.host_orient(rp.numpy(), ci.numpy())
    capi.set_option("tc.algo", "bs")
    with capi.DeviceGraph(orp, oci, md) as g:
        want = g.tc()
    capi.set_option("tc.algo", "rank")
    nv = len(orp) - 1
    for flat in (5, 1):
        capi.set_option("tc.flat", flat)
        with capi.DeviceGraph(orp, oci, md) as g:
            assert g.tc() == want, flat
    capi.set_option("tc.flat", 5)
    for shard in ("source", "dest"):
        capi.set_option("tc.shard", shard)
        total = 0
        for lo, hi in ((0, nv // 3), (nv // 3, nv // 2), (nv // 2, nv)):
            with capi.DeviceGraph(orp, oci, md) as g:
                g.set_source_range(lo, hi)
                total += g.tc()
        assert total == want, shard


def test_mapped_graph_feeds_the_host_entry_points(tmp_path):
    """gm_host_map_graph -> gm_tc_host: the read-only, driver-registered mapping is uploaded like pinned memory"""
    rp, ci = _graph("rmat14")
    orp, oci, md = _dag(rp, ci)
    prefix = str(tmp_path / "dag")
    capi.write_graph(prefix, orp, oci, md)
    m_rp, m_ci, m_md, pinned = capi.map_graph(prefix)
    assert isinstance(pinned, bool)                       # registration needs read-only host-register support
    assert capi.tc_host(m_rp, m_ci, m_md) == GOLD["rmat14"]["tc"]
    u_rp, u_ci, u_md, _ = capi.map_graph(prefix, pin=False)
    assert capi.tc_host(u_rp, u_ci, u_md) == GOLD["rmat14"]["tc"]


@pytest.mark.parametrize("name", ["rmat8", "rmat12", "rmat14", "shaped3000"])
def test_device_orientation_matches_the_host_and_the_oracle(name):
    """k_orient (support.cu) against Graph::orientation as restated by the host library and by the oracle:
    the same rows in the same order, bit for bit"""
    rp, ci = _graph(name)
    want_rp, want_ci, _ = capi.host_orient(rp, ci)
    o_rp, o_ci = oracle.orient(rp, ci)[:2]
    assert np.array_equal(want_rp, o_rp) and np.array_equal(want_ci, o_ci)
    with capi.DeviceGraph(rp, ci, 0) as g:
        got_rp, got_ci = g.orient()
        assert np.array_equal(got_rp, want_rp) and np.array_equal(got_ci, want_ci)
        back_rp, back_ci = g.download()
        assert np.array_equal(back_rp, rp) and np.array_equal(back_ci, ci)


@pytest.mark.parametrize("name", ["rmat10", "rmat14", "shaped3000"])
def test_device_partition_matches_the_host(name):
    """gm_graph_partition (1-hop induced part built on the device) against gm_host_partition_part: same vertex
    set, relabelling, rows and local range; and the part counts its share of the triangles (the reference's
    multi-GPU TC scheme, triangle/multigpu.cu:45-75)"""
    rp, ci = _graph(name)
    orp, oci, md = _dag(rp, ci)
    nv = len(orp) - 1
    total = 0
    with capi.DeviceGraph(orp, oci, md) as g:
        for lo, hi in ((0, nv // 3), (nv // 3, nv // 3), (nv // 3, nv)):
            want = capi.host_partition_part(orp, oci, lo, hi)
            part, idx, lb, le = g.partition(lo, hi)
            with part:
                got_rp, got_ci = part.download()
                assert np.array_equal(got_rp, want[0]) and np.array_equal(got_ci, want[1])
                assert np.array_equal(idx, want[2]) and (lb, le) == (want[3], want[4])
                part.set_source_range(lb, le)
                total += part.tc()
    assert total == GOLD[name]["tc"]
