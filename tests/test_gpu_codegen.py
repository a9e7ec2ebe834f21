"""Kernels emitted by graphminer_b200/codegen.py, built with nvcc on this box and run through
gm_graph_device_view on a DeviceGraph: bit-exact against the oracle's sgl / motif counts, the reference OMP goldens
and the built-in solvers."""
import json
import os

import numpy as np
import pytest

import oracle
from graphminer_b200 import capi, codegen
from graphminer_b200.rmat import rmat_graph, shaped_graph

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rmat_counts.json")))


@pytest.fixture(scope="module")
def graphs():
    return {"rmat10": tuple(t.numpy() for t in rmat_graph(10)), "shaped3000": tuple(t.numpy() for t in shaped_graph(3000, 40000, 0x5EED004C))}


@pytest.mark.parametrize("name", ["triangle", "diamond", "rectangle", "house", "pentagon", "clique4", "clique5"])
def test_edge_induced_patterns(name, graphs, tmp_path):
    kern = codegen.compile(name, induced=False, build_dir=str(tmp_path))
    for gname, (rp, ci) in graphs.items():
        gold = GOLD[gname]
        want = {"triangle": gold["tc"], "diamond": gold["diamond"], "rectangle": gold["rectangle"], "house": gold["house"],
                "pentagon": gold["pentagon"], "clique4": gold["clique4"], "clique5": gold["clique5"]}[name]
        with capi.DeviceGraph(rp, ci, 0) as g:
            assert kern.count(g) == want, (name, gname)
            assert kern.count(g) == want                      # cached task list


@pytest.mark.parametrize("name,slot", [("star3", 0), ("path4", 1), ("tailed_triangle", 2), ("rectangle", 3), ("diamond", 4), ("clique4", 5)])
def test_vertex_induced_patterns_are_the_4_motifs(name, slot, graphs, tmp_path):
    kern = codegen.compile(name, induced=True, build_dir=str(tmp_path))
    for gname, (rp, ci) in graphs.items():
        with capi.DeviceGraph(rp, ci, 0) as g:
            assert kern.count(g) == GOLD[gname]["motif4"][slot], (name, gname)


def test_user_defined_pattern(graphs, tmp_path):
    """a pattern nobody hand-wrote: the bull (triangle with two horns), against the host interpretation of its plan"""
    bull = codegen.Pattern(5, [(0, 1), (0, 2), (1, 2), (1, 3), (2, 4)], "bull")
    rp, ci = (t.numpy() for t in rmat_graph(8))
    for induced in (False, True):
        kern = codegen.compile(bull, induced=induced, build_dir=str(tmp_path))
        with capi.DeviceGraph(rp, ci, 0) as g:
            assert kern.count(g) == codegen.count_on_host(bull, rp, ci, induced=induced)
