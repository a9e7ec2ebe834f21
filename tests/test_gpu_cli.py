"""The CLI drop-ins (bin/*_gpu_base, *_multigpu) print the reference's result lines
(src/{triangle,clique,sgl,motif}/main.cc) with the known-answer counts."""
import json
import os
import re
import subprocess

import pytest

from graphminer_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))["citeseer"]


@pytest.fixture(scope="module")
def graph_prefix(tmp_path_factory, citeseer):
    rp, ci, md = citeseer
    d = tmp_path_factory.mktemp("citeseer")
    prefix = str(d / "graph")
    capi.write_graph(prefix, rp, ci, md)
    return prefix


def run(binary, *args):
    exe = os.path.join(BIN, binary)
    if not os.path.exists(exe):
        pytest.skip(f"{exe} not built (make apps)")
    return subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True, timeout=300, check=True).stdout


def test_tc(graph_prefix):
    for b in ("tc_gpu_base", "tc_multigpu"):
        out = run(b, graph_prefix)
        assert f"total_num_triangles = {KAT['tc']}\n" in out
        assert "runtime [gpu_base] = " in out and "Traversed Edges Per Second (TEPS)" in out


def test_clique(graph_prefix):
    for b in ("clique_gpu_base", "kcl_gpu_base", "clique_multigpu"):
        assert f"num_4-cliques = {KAT['clique4']}\n" in run(b, graph_prefix, 4)
    assert f"num_5-cliques = {KAT['clique5']}\n" in run("clique_gpu_base", graph_prefix, 5)
    out = run("clique_gpu_base", graph_prefix, 12)
    assert "Not supported right now" in out and "num_12-cliques = 0" in out


def test_sgl(graph_prefix):
    for p in ("diamond", "rectangle", "house", "pentagon"):
        assert f"total_num = {KAT[p]}\n" in run("sgl_gpu_base", graph_prefix, p)
    assert f"total_num = {KAT['diamond']}\n" in run("sgl_multigpu", graph_prefix, "diamond", 2)
    assert "Not implemented" in run("sgl_gpu_base", graph_prefix, "dumbbell")


def test_motif(graph_prefix):
    for b in ("motif_gpu_base", "motif_gpu_formula", "motif_multigpu"):
        out = run(b, graph_prefix, 4)
        got = [int(x) for x in re.findall(r"pattern \d+: (\d+)", out)]
        assert got == KAT["motif4"], b
    out = run("motif_gpu_base", graph_prefix, 3)
    assert [int(x) for x in re.findall(r"pattern \d+: (\d+)", out)] == KAT["motif3"]


def test_usage_exit_code():
    exe = os.path.join(BIN, "tc_gpu_base")
    if not os.path.exists(exe):
        pytest.skip("not built")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stdout


def test_tc_unsorted_input(tmp_path_factory, citeseer):
    """`tc <graph> 1 1024 0`: adj_sorted = 0 sorts the neighbour lists after loading (triangle/main.cc:21-22)."""
    import numpy as np
    rp, ci, md = citeseer
    rng = np.random.default_rng(1)
    shuffled = ci.copy()
    for v in range(len(rp) - 1):
        shuffled[rp[v]:rp[v + 1]] = rng.permutation(ci[rp[v]:rp[v + 1]])
    assert not capi.check_sorted(rp, shuffled)
    prefix = str(tmp_path_factory.mktemp("citeseer_unsorted") / "graph")
    capi.write_graph(prefix, rp, shuffled, md)
    out = run("tc_gpu_base", prefix, 1, 1024, 0)
    assert "Sorting the neighbor lists" in out and f"total_num_triangles = {KAT['tc']}\n" in out
