"""Deterministic synthetic graph generators (R-MAT and "shaped" R-MAT).

The reference ships no generator; its benchmarks read graphs in its own binary
CSR format (`graph.meta.txt` / `graph.vertex.bin` / `graph.edge.bin`,
/root/reference/src/common/graph.cc:19-41, README.md:82-100).  SURVEY.md §8(d)
fixes the synthetic workloads for this repo: Graph500 R-MAT
(a,b,c,d)=(0.57,0.19,0.19,0.05), edge factor 16, a fixed seed per scale, a
random vertex-id permutation, self-loops dropped, symmetrised, sorted and
de-duplicated.

Everything here is pure 64-bit integer arithmetic on torch tensors, so the same
code yields bit-identical graphs on CPU (tests, the CPU oracle) and on CUDA
(bench.py generates scale-22+ graphs directly in HBM).  torch is plumbing only.
"""
from __future__ import annotations

import torch

_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB


def _s64(x: int) -> int:
    """Python int -> the signed 64-bit value with the same bit pattern."""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    """Logical shift right on int64 tensors (torch's >> is arithmetic)."""
    return (x >> k) & ((1 << (64 - k)) - 1)


def _mix64(x: torch.Tensor) -> torch.Tensor:
    """splitmix64 finaliser; int64 tensors wrap modulo 2^64 like uint64."""
    x = x ^ _lsr(x, 30)
    x = x * _s64(_C1)
    x = x ^ _lsr(x, 27)
    x = x * _s64(_C2)
    x = x ^ _lsr(x, 31)
    return x


def _permute_ids(x: torch.Tensor, bits: int, seed: int) -> torch.Tensor:
    """A seed-dependent bijection on [0, 2^bits): odd multiplies and xor-shifts."""
    mask = (1 << bits) - 1
    k1 = (_mix64_int(seed ^ 0xA5A5A5A5) | 1) & mask
    k2 = (_mix64_int(seed ^ 0x5A5A5A5A) | 1) & mask
    c1 = _mix64_int(seed ^ 0x1234567) & mask
    h = max(1, bits // 2)
    x = (x * k1 + c1) & mask
    x = x ^ (x >> h)
    x = (x * k2) & mask
    x = x ^ (x >> h)
    x = (x * k1) & mask
    x = x ^ (x >> h)
    return x


def _mix64_int(x: int) -> int:
    x &= _M64
    x ^= x >> 30
    x = (x * _C1) & _M64
    x ^= x >> 27
    x = (x * _C2) & _M64
    x ^= x >> 31
    return x


def rmat_edges(scale: int, n_samples: int, seed: int, probs=(0.57, 0.19, 0.19, 0.05),
               device="cpu", attempt: int = 0, index_offset: int = 0,
               index: torch.Tensor | None = None):
    """Sample `n_samples` directed R-MAT pairs over 2^scale ids (before permutation).

    Edge i, level l use 16 bits of mix64(seed, attempt, i, l//4); thresholds are
    the cumulative probabilities scaled to 2^16.
    """
    a, b, c, _ = probs
    ta = int(a * 65536)
    tab = int((a + b) * 65536)
    tabc = int((a + b + c) * 65536)
    if index is None:
        index = torch.arange(index_offset, index_offset + n_samples, dtype=torch.int64, device=device)
    base = index * _s64(_GOLD) + _s64(_mix64_int(seed * 0x100000001B3 + attempt * 0x51ED27))
    src = torch.zeros_like(index)
    dst = torch.zeros_like(index)
    h = None
    for lvl in range(scale):
        if lvl % 4 == 0:
            h = _mix64(base + _s64((lvl // 4 + 1) * 0xD6E8FEB86659FD93))
        r = _lsr(h, 16 * (lvl % 4)) & 0xFFFF
        sbit = (r >= tab).to(torch.int64)                         # quadrants c,d -> src bit 1
        dbit = ((r >= ta) & (r < tab) | (r >= tabc)).to(torch.int64)  # quadrants b,d -> dst bit 1
        src = (src << 1) | sbit
        dst = (dst << 1) | dbit
    return src, dst


def edges_to_csr(src: torch.Tensor, dst: torch.Tensor, nv: int):
    """Drop self-loops, symmetrise, sort, dedupe -> (rowptr int64[nv+1], colidx int32[ne])."""
    keep = src != dst
    src, dst = src[keep], dst[keep]
    key = torch.cat([(src << 32) | dst, (dst << 32) | src])
    del src, dst
    key = torch.unique(key, sorted=True)
    row = key >> 32
    col = (key & 0xFFFFFFFF).to(torch.int32)
    del key
    counts = torch.bincount(row, minlength=nv)
    rowptr = torch.zeros(nv + 1, dtype=torch.int64, device=col.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr, col


def rmat_graph(scale: int, edge_factor: int = 16, seed: int | None = None,
               probs=(0.57, 0.19, 0.19, 0.05), device="cpu", chunk: int = 1 << 26):
    """Undirected R-MAT graph of 2^scale vertices in CSR."""
    if seed is None:
        seed = 0x5EED0000 + scale
    nv = 1 << scale
    m = edge_factor * nv
    srcs, dsts = [], []
    for off in range(0, m, chunk):
        n = min(chunk, m - off)
        s, d = rmat_edges(scale, n, seed, probs, device, index_offset=off)
        srcs.append(_permute_ids(s, scale, seed))
        dsts.append(_permute_ids(d, scale, seed))
    src = torch.cat(srcs) if len(srcs) > 1 else srcs[0]
    dst = torch.cat(dsts) if len(dsts) > 1 else dsts[0]
    del srcs, dsts
    return edges_to_csr(src, dst, nv)


def shaped_graph(nv: int, n_samples: int, seed: int, probs=(0.57, 0.19, 0.19, 0.05),
                 device="cpu", max_rounds: int = 64):
    """R-MAT over the next power of two with rejection of ids >= nv (configs 4/5).

    A sample whose permuted endpoints fall outside [0, nv) is re-drawn with the
    next `attempt` counter until it lands inside (expected acceptance (nv/2^b)^2).
    """
    bits = max(1, (nv - 1).bit_length())
    index = torch.arange(n_samples, dtype=torch.int64, device=device)
    out_s = torch.empty(n_samples, dtype=torch.int64, device=device)
    out_d = torch.empty(n_samples, dtype=torch.int64, device=device)
    pending = index
    for attempt in range(max_rounds):
        if pending.numel() == 0:
            break
        s, d = rmat_edges(bits, pending.numel(), seed, probs, device, attempt=attempt, index=pending)
        s = _permute_ids(s, bits, seed)
        d = _permute_ids(d, bits, seed)
        ok = (s < nv) & (d < nv)
        out_s[pending[ok]] = s[ok]
        out_d[pending[ok]] = d[ok]
        pending = pending[~ok]
    if pending.numel():
        # vanishingly unlikely after 64 rounds; make those samples self-loops (dropped later)
        out_s[pending] = 0
        out_d[pending] = 0
    return edges_to_csr(out_s, out_d, nv)


def orient_dag(rowptr: torch.Tensor, colidx: torch.Tensor):
    """Degree/id orientation, same rule as Graph::orientation (graph.cc:233-279).

    Keep u->v iff deg(v) > deg(u) or (deg(v) == deg(u) and v > u); order-preserving filter.
    Torch mirror used by bench.py to build device-resident inputs; the product's
    host implementation is gm_host_orient in the C-ABI library.
    """
    nv = rowptr.numel() - 1
    deg = rowptr[1:] - rowptr[:-1]
    src = torch.repeat_interleave(torch.arange(nv, dtype=torch.int64, device=rowptr.device), deg)
    dl = colidx.long()
    keep = (deg[dl] > deg[src]) | ((deg[dl] == deg[src]) & (dl > src))
    new_deg = torch.bincount(src[keep], minlength=nv)
    new_rowptr = torch.zeros(nv + 1, dtype=torch.int64, device=rowptr.device)
    torch.cumsum(new_deg, 0, out=new_rowptr[1:])
    return new_rowptr, colidx[keep]
