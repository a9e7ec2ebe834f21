"""Pattern -> loop-nest code generator over the gm operator API.

The reference ships a Python prototype (codegen/vertex_gen.py, hybrid_gen.py) that turns a small pattern graph
into an AutoMine-style loop nest: a *matching order* (one pattern vertex per loop level, AutoMine) and a
*symmetry order* (id(v_i) < id(v_j) restrictions that break the pattern's automorphisms, GraphZero / GraphPi),
printed as C++ over its VertexSet operators.  This module does the same for the B200 engine and goes one step
further: it emits a complete CUDA translation unit -- a warp-per-edge DFS kernel written against
`include/gm/set_ops.cuh` / `gm/graph_gpu.cuh` plus a C entry point -- that `compile()` builds with nvcc for
sm_100a and loads, so a user-defined pattern runs on a `capi.DeviceGraph` like the built-in solvers:

    from graphminer_b200 import codegen, capi
    house = codegen.Pattern(5, [(0, 1), (0, 2), (1, 2), (1, 3), (0, 4), (3, 4)])
    kern = codegen.compile(house)                      # edge-induced (sgl semantics); induced=True: motif semantics
    with capi.DeviceGraph(rowptr, colidx) as g:
        count = kern.count(g)

Schedule of the generated kernel (one warp per task = one data edge matched to the first pattern edge):
  * level i >= 2 draws v_i from  S_i = ∩_{j<i, (j,i) in P} N(v_j)  [ \\ ∪_{j<i, (j,i) not in P} N(v_j)  if induced ],
    restricted to ids below every earlier vertex the symmetry order puts above v_i, and different from the
    earlier vertices;
  * S_i is refined INCREMENTALLY: as soon as v_t is bound, every later level's partial set absorbs N(v_t)
    (intersect / difference_set into a per-warp scratch slice), so work shared by sibling subtrees is done once
    -- AutoMine's schedule;  a set with a single operand so far is a view of the CSR row (no copy);
  * the last level only counts (intersect_num / difference_num with bound, or count_smaller on a materialised set).
Bit-exactness is checked against the oracle's sgl / motif counts in tests/test_gpu_codegen.py.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import itertools
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


class Pattern:
    """An undirected pattern graph on vertices 0..n-1 (connected, 3 <= n <= 6)."""

    def __init__(self, n, edges, name=None):
        self.n = int(n)
        self.edges = sorted({(min(a, b), max(a, b)) for a, b in edges if a != b})
        self.name = name or "p%d_%s" % (n, hashlib.sha1(repr(self.edges).encode()).hexdigest()[:8])
        if not (3 <= self.n <= 6):
            raise ValueError("patterns of 3..6 vertices")
        if any(not (0 <= a < n and 0 <= b < n) for a, b in self.edges):
            raise ValueError("edge end point out of range")
        adj = self.adjacency()
        seen, stack = {0}, [0]
        while stack:
            for w in adj[stack.pop()]:
                if w not in seen:
                    seen.add(w); stack.append(w)
        if len(seen) != self.n:
            raise ValueError("pattern must be connected")

    def adjacency(self):
        adj = [set() for _ in range(self.n)]
        for a, b in self.edges:
            adj[a].add(b); adj[b].add(a)
        return adj

    def has_edge(self, a, b):
        return (min(a, b), max(a, b)) in self.edges

    def automorphisms(self):
        es = set(self.edges)
        out = []
        for perm in itertools.permutations(range(self.n)):
            if all((min(perm[a], perm[b]), max(perm[a], perm[b])) in es for a, b in self.edges):
                out.append(perm)
        return out


NAMED = {
    "triangle": Pattern(3, [(0, 1), (0, 2), (1, 2)], "triangle"),
    "wedge": Pattern(3, [(0, 1), (0, 2)], "wedge"),
    "diamond": Pattern(4, [(0, 1), (0, 2), (1, 2), (0, 3), (1, 3)], "diamond"),
    "rectangle": Pattern(4, [(0, 1), (1, 2), (2, 3), (0, 3)], "rectangle"),
    "clique4": Pattern(4, [(a, b) for a in range(4) for b in range(a)], "clique4"),
    "tailed_triangle": Pattern(4, [(0, 1), (0, 2), (1, 2), (0, 3)], "tailed_triangle"),
    "star3": Pattern(4, [(0, 1), (0, 2), (0, 3)], "star3"),
    "path4": Pattern(4, [(0, 1), (1, 2), (2, 3)], "path4"),
    "house": Pattern(5, [(0, 1), (0, 2), (1, 2), (1, 3), (0, 4), (3, 4)], "house"),
    "pentagon": Pattern(5, [(0, 1), (1, 2), (2, 3), (3, 4), (0, 4)], "pentagon"),
    "clique5": Pattern(5, [(a, b) for a in range(5) for b in range(a)], "clique5"),
}


# ---- matching order (AutoMine) and symmetry order (GraphZero) ------------------------------------------------
def matching_order(p: Pattern):
    """Greedy: start at the densest edge, then always the vertex with most neighbours among the matched ones
    (ties: higher pattern degree, lower id) -- every prefix is connected and intersections come early."""
    adj = p.adjacency()
    deg = [len(a) for a in adj]
    a0, b0 = max(p.edges, key=lambda e: (deg[e[0]] + deg[e[1]], len(adj[e[0]] & adj[e[1]]), -e[0], -e[1]))
    first = (a0, b0) if (deg[a0], -a0) >= (deg[b0], -b0) else (b0, a0)
    order = list(first)
    while len(order) < p.n:
        rest = [v for v in range(p.n) if v not in order]
        order.append(max(rest, key=lambda v: (len(adj[v] & set(order)), deg[v], -v)))
    return order


def symmetry_order(p: Pattern, order):
    """Restrictions (i, j), i > j as LEVELS of `order`, meaning id(v_i) < id(v_j): for the first level whose vertex
    some remaining automorphism moves, every other vertex of its orbit must be smaller; continue in the stabiliser.
    Exactly one member of every automorphism class of embeddings satisfies all restrictions (GraphZero)."""
    pos = {v: i for i, v in enumerate(order)}
    group = p.automorphisms()
    cons = []
    for lvl, v in enumerate(order):
        orbit = {g[v] for g in group}
        for u in sorted(orbit - {v}, key=lambda x: pos[x]):
            assert pos[u] > lvl, "orbit members of an unfixed vertex come later in the matching order"
            cons.append((pos[u], lvl))
        group = [g for g in group if g[v] == v]
        if len(group) == 1:
            break
    return cons


def plan(p: Pattern):
    """(order, conn, disc, upper): the loop-nest plan generate() emits, as data"""
    order = matching_order(p)
    cons = symmetry_order(p, order)
    n = p.n
    conn = [[j for j in range(i) if p.has_edge(order[i], order[j])] for i in range(n)]
    disc = [[j for j in range(i) if not p.has_edge(order[i], order[j])] for i in range(n)]
    upper = [[j for (i2, j) in cons if i2 == i] for i in range(n)]
    return order, conn, disc, upper


def count_on_host(p: Pattern, rowptr, colidx, induced=False) -> int:
    """The same plan interpreted with Python sets (small graphs only): what the generated kernel must return."""
    order, conn, disc, upper = plan(p)
    nv = len(rowptr) - 1
    adj = [set(int(x) for x in colidx[rowptr[v]:rowptr[v + 1]]) for v in range(nv)]
    n = p.n

    def rec(i, vs):
        if i == n:
            return 1
        cand = None
        for j in conn[i]:
            cand = set(adj[vs[j]]) if cand is None else cand & adj[vs[j]]
        if induced:
            for j in disc[i]:
                cand = cand - adj[vs[j]]
        cand = cand - set(vs)
        if upper[i]:
            ub = min(vs[j] for j in upper[i])
            cand = {x for x in cand if x < ub}
        return sum(rec(i + 1, vs + [x]) for x in cand)

    total = 0
    for v0 in range(nv):
        for v1 in adj[v0]:
            if 0 in upper[1] and not v1 < v0:
                continue
            total += rec(2, [v0, v1])
    return total


# ---- code emission -----------------------------------------------------------------------------------------
class _Set:
    """symbolic vertex set of a later level during emission: a CSR row view or a materialised scratch buffer"""

    def __init__(self, ptr, size):
        self.ptr, self.size = ptr, size


def generate(p: Pattern, induced: bool = False) -> str:
    order = matching_order(p)
    cons = symmetry_order(p, order)
    n = p.n
    conn = [[j for j in range(i) if p.has_edge(order[i], order[j])] for i in range(n)]
    disc = [[j for j in range(i) if not p.has_edge(order[i], order[j])] for i in range(n)]
    upper = [[j for (i2, j) in cons if i2 == i] for i in range(n)]          # id(v_i) < id(v_j)
    assert conn[1] == [0]
    sym_break = 1 if 0 in upper[1] else 0
    last = n - 1
    L = []
    emit = L.append
    emit("// GENERATED by graphminer_b200/codegen.py -- pattern %s, %s-induced" % (p.name, "vertex" if induced else "edge"))
    emit("// matching order (pattern vertices): %s; symmetry order: %s" %
         (order, ", ".join("v%d < v%d" % c for c in cons) or "none"))
    emit('#include <cstdint>\n#include <cstdio>\n#include <algorithm>\n#include <cuda_runtime.h>')
    emit('#include "gm/graph_gpu.cuh"\n#include "gminer_b200.h"\nusing namespace gm;')
    emit("constexpr int kBufs = %d;" % (n * n))
    emit("__global__ void __launch_bounds__(256) pattern_kernel(GraphGPU g, vidType *scratch, int64_t max_deg, unsigned long long *ticket, AccType *total) {")
    emit("  const int lane = threadIdx.x & 31;")
    emit("  vidType *buf = scratch + ((int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5) * (max_deg * kBufs);")
    emit("  AccType cnt = 0;")
    emit("  eidType cur = 0, end = 0;")
    emit("  while (true) {")
    emit("    if (cur >= end) {")
    emit("      unsigned long long t = 0; if (lane == 0) t = atomicAdd(ticket, 4ull); t = __shfl_sync(kFullMask, t, 0);")
    emit("      cur = eidType(t); end = min(cur + 4, g.num_tasks); if (cur >= g.num_tasks) break;")
    emit("    }")
    emit("    const eidType e = cur++;")
    emit("    const vidType v0 = g.get_src(e), v1 = g.get_dst(e);")
    if 0 in upper[1] and not sym_break:
        raise AssertionError
    state = [None] * n          # partial candidate set of every later level: (ptr expr, size expr) or None

    def refine(t, ind):
        """v_t has just been bound: every later level absorbs N(v_t).  The LAST level is left alone at t = n-2: its final
        operand is folded into the counting operator."""
        for i in range(max(t + 1, 2), n):
            if i == last and t == last - 1:
                continue
            dst = "buf + max_deg * %d" % (i * n + t)
            if t in conn[i]:
                if state[i] is None:
                    state[i] = ("g.N(v%d)" % t, "g.get_degree(v%d)" % t)
                    if induced:                       # earlier non-neighbours had nothing to be subtracted from yet
                        for j in [x for x in disc[i] if x < t]:
                            d2 = "buf + max_deg * %d" % (i * n + j)
                            emit(ind + "const vidType n%d_%d = difference_set_except(%s, vidType(%s), g.N(v%d), g.get_degree(v%d), vidType(kVidMax), v%d, %s);" %
                                 (i, j, state[i][0], state[i][1], j, j, j, d2))
                            state[i] = (d2, "n%d_%d" % (i, j))
                else:
                    emit(ind + "const vidType n%d_%d = intersect(%s, vidType(%s), g.N(v%d), g.get_degree(v%d), %s);" %
                         (i, t, state[i][0], state[i][1], t, t, dst))
                    state[i] = (dst, "n%d_%d" % (i, t))
            elif induced and t in disc[i] and state[i] is not None:
                emit(ind + "const vidType n%d_%d = difference_set_except(%s, vidType(%s), g.N(v%d), g.get_degree(v%d), vidType(kVidMax), v%d, %s);" %
                     (i, t, state[i][0], state[i][1], t, t, t, dst))
                state[i] = (dst, "n%d_%d" % (i, t))

    ind = "    "
    stack = []
    refine(0, ind)
    refine(1, ind)
    for i in range(2, n):
        ub = None
        if upper[i]:
            ub = "ub%d" % i
            emit(ind + "const vidType %s = %s;" % (ub, _min_expr(["v%d" % j for j in upper[i]])))
        if i < last:
            s = state[i]
            assert s is not None
            emit(ind + "for (vidType i%d = 0; i%d < vidType(%s); i%d++) {" % (i, i, s[1], i))
            ind += "  "
            emit(ind + "const vidType v%d = (%s)[i%d];" % (i, s[0], i))
            if ub:
                emit(ind + "if (v%d >= %s) break;" % (i, ub))
            if not induced:
                for j in disc[i]:
                    emit(ind + "if (v%d == v%d) continue;" % (i, j))
            stack.append(list(state))
            refine(i, ind)
            continue
        # ---- last level: count; the operand of level n-2 (if any) is folded into the counting operator ----
        t = last - 1
        s = state[last]
        ubx = ub or "vidType(kVidMax)"
        anc = [] if induced else list(disc[last])              # earlier vertices a candidate must differ from
        if t in conn[last] and s is None:                       # the only operand: a CSR row
            pre_disc = [j for j in disc[last] if j < t] if induced else []
            if pre_disc:                                        # induced: subtract the earlier non-neighbours' rows first
                ptr, size = "g.N(v%d)" % t, "g.get_degree(v%d)" % t
                for k, j in enumerate(pre_disc):
                    if k == len(pre_disc) - 1:
                        emit(ind + "cnt += difference_num_except(%s, vidType(%s), g.N(v%d), g.get_degree(v%d), %s, v%d);" % (ptr, size, j, j, ubx, j))
                    else:
                        d2 = "buf + max_deg * %d" % (last * n + j)
                        emit(ind + "const vidType n%d_%d = difference_set_except(%s, vidType(%s), g.N(v%d), g.get_degree(v%d), vidType(kVidMax), v%d, %s);" %
                             (last, j, ptr, size, j, j, j, d2))
                        ptr, size = d2, "n%d_%d" % (last, j)
            else:
                emit(ind + "{ AccType c = count_smaller(%s, g.N(v%d), g.get_degree(v%d));" % (ubx, t, t))
                for j in anc:
                    emit(ind + "  if (v%d < %s && binary_search(g.N(v%d), v%d, g.get_degree(v%d))) c--;" % (j, ubx, t, j, t))
                emit(ind + "  if (lane == 0) cnt += c; }")
        elif t in conn[last]:
            if anc:
                emit(ind + "{ const vidType anc[%d] = {%s};" % (len(anc), ", ".join("v%d" % j for j in anc)))
                emit(ind + "  cnt += intersect_num(%s, vidType(%s), g.N(v%d), g.get_degree(v%d), %s, anc, %d); }" % (s[0], s[1], t, t, ubx, len(anc)))
            else:
                emit(ind + "cnt += intersect_num(%s, vidType(%s), g.N(v%d), g.get_degree(v%d), %s);" % (s[0], s[1], t, t, ubx))
        elif induced and t in disc[last]:
            assert s is not None
            emit(ind + "cnt += difference_num_except(%s, vidType(%s), g.N(v%d), g.get_degree(v%d), %s, v%d);" % (s[0], s[1], t, t, ubx, t))
        else:                                                   # v_{n-2} does not touch the last level (edge-induced only)
            assert s is not None
            emit(ind + "{ AccType c = count_smaller(%s, %s, vidType(%s));" % (ubx, s[0], s[1]))
            for j in anc:
                emit(ind + "  if (v%d < %s && binary_search(%s, v%d, vidType(%s))) c--;" % (j, ubx, s[0], j, s[1]))
            emit(ind + "  if (lane == 0) cnt += c; }")
    for i in range(last - 1, 1, -1):
        ind = ind[:-2]
        emit(ind + "}")
        state[:] = stack.pop()
    emit("  }")
    emit("  cnt = warp_reduce(cnt);")
    emit("  if (lane == 0 && cnt) atomicAdd(total, cnt);")
    emit("}")
    emit(_HOST % {"sym_break": sym_break, "name": p.name})
    return "\n".join(L)


def _min_expr(xs):
    e = xs[0]
    for x in xs[1:]:
        e = "min(%s, %s)" % (e, x)
    return e


_HOST = r'''
extern "C" int gm_pattern_count(gm_graph_t *g, uint64_t *count) {
  GraphGPU view; void *stream = nullptr; int sms = 0, max_deg = 0;
  int rc = gm_graph_device_view(g, %(sym_break)d, &view, sizeof view, &stream, &sms, &max_deg);
  if (rc != GM_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t md = std::max(max_deg, 1);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pattern_kernel, 256, 0);
  int64_t blocks = std::min<int64_t>(std::max<int64_t>((view.num_tasks + 31) / 32, 1), int64_t(std::max(occ, 1)) * sms);
  size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
  const int64_t per_block = md * kBufs * 8 * int64_t(sizeof(vidType));
  blocks = std::max<int64_t>(1, std::min<int64_t>(blocks, int64_t(double(free_b) * 0.6) / per_block));
  vidType *scratch = nullptr; unsigned long long *ctr = nullptr, h[2] = {0, 0};
  if (cudaMallocAsync(reinterpret_cast<void **>(&scratch), size_t(blocks) * size_t(per_block), s) != cudaSuccess ||
      cudaMallocAsync(reinterpret_cast<void **>(&ctr), 2 * sizeof(unsigned long long), s) != cudaSuccess) { cudaGetLastError(); return GM_ENOMEM; }
  cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), s);
  if (view.num_tasks > 0) pattern_kernel<<<unsigned(blocks), 256, 0, s>>>(view, scratch, md, ctr, ctr + 1);
  cudaMemcpyAsync(h, ctr, sizeof h, cudaMemcpyDeviceToHost, s);
  cudaError_t e = cudaStreamSynchronize(s);
  cudaFreeAsync(scratch, s); cudaFreeAsync(ctr, s);
  if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { fprintf(stderr, "generated kernel %(name)s: %%s\n", cudaGetErrorString(e)); return GM_ECUDA; }
  *count = h[1];
  return GM_OK;
}
'''


# ---- build + load ------------------------------------------------------------------------------------------
class CompiledPattern:
    def __init__(self, lib_path, pattern, induced, source):
        self.path, self.pattern, self.induced, self.source = lib_path, pattern, induced, source
        self._lib = C.CDLL(lib_path)
        self._lib.gm_pattern_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]

    def count(self, device_graph) -> int:
        from . import capi
        out = C.c_uint64(0)
        capi.check(self._lib.gm_pattern_count(device_graph._h, C.byref(out)))
        return out.value


def compile(pattern, induced=False, build_dir=None, nvcc="/usr/local/cuda/bin/nvcc") -> CompiledPattern:  # noqa: A001
    if isinstance(pattern, str):
        pattern = NAMED[pattern]
    src = generate(pattern, induced)
    build_dir = build_dir or os.path.join(tempfile.gettempdir(), "gm_codegen")
    os.makedirs(build_dir, exist_ok=True)
    tag = "%s_%s_%s" % (pattern.name, "vi" if induced else "ei", hashlib.sha1(src.encode()).hexdigest()[:10])
    cu, so = os.path.join(build_dir, tag + ".cu"), os.path.join(build_dir, tag + ".so")
    if not os.path.exists(so):
        with open(cu, "w") as f:
            f.write(src)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared", "-Xcompiler", "-fPIC",
               "-I" + os.path.join(ROOT, "include"), cu, "-o", so, "-L" + HERE, "-lgminer_b200", "-Xlinker", "-rpath", "-Xlinker", HERE]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed on the generated kernel:\n" + r.stderr[-3000:] + "\n--- source ---\n" + src)
    return CompiledPattern(so, pattern, induced, src)


if __name__ == "__main__":
    import sys
    name = sys.argv[1] if len(sys.argv) > 1 else "house"
    print(generate(NAMED[name], induced="--induced" in sys.argv))
