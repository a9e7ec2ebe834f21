"""Every device set operator (gm/set_ops.cuh via gm_intersect_batch) and every streaming variant
(bsearch / merge+TMA / hash / gallop) against the CPU oracle on seeded random sorted lists:
empty, single-element, disjoint, identical, size ratios 1:1 .. 1:10^4, bounds at every position class."""
import numpy as np
import pytest

import oracle
from graphminer_b200 import capi

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def make_pairs(seed=11, align=False):
    rng = np.random.default_rng(seed)
    pairs = [(np.array([], np.int32), np.array([], np.int32)),
             (np.array([5], np.int32), np.array([], np.int32)),
             (np.array([], np.int32), np.array([1, 2, 3], np.int32)),
             (np.array([3], np.int32), np.array([3], np.int32)),
             (np.array([3], np.int32), np.array([4], np.int32)),
             (np.arange(0, 100, 2, dtype=np.int32), np.arange(1, 101, 2, dtype=np.int32)),
             (np.arange(40, dtype=np.int32), np.arange(40, dtype=np.int32)),
             (np.arange(33, dtype=np.int32), np.arange(31, 64, dtype=np.int32)),
             (np.arange(1000, dtype=np.int32), np.arange(500, 3000, 3, dtype=np.int32))]
    shapes = [(1, 1), (2, 31), (32, 32), (33, 64), (5, 1000), (1000, 5), (100, 100), (257, 300), (700, 900),
              (3, 30000), (1200, 1400), (2500, 2500), (64, 4096), (1, 5000)]
    for rep in range(4):
        for na, nb in shapes:
            hi = int(max(na, nb) * rng.choice([1.2, 2.0, 10.0])) + 2
            a = np.unique(rng.integers(0, hi, na)).astype(np.int32)
            b = np.unique(rng.integers(0, hi, nb)).astype(np.int32)
            pairs.append((a, b))
    return pairs


def pack(pairs, seed=3):
    """Lay the lists out in one pool at arbitrary (unaligned) offsets with junk between them."""
    rng = np.random.default_rng(seed)
    chunks, a_off, b_off, pos = [], [], [], 0
    for a, b in pairs:
        for lst, offs in ((a, a_off), (b, b_off)):
            gap = int(rng.integers(0, 4))
            chunks.append(np.full(gap, -7, np.int32)); pos += gap
            offs.append(pos); chunks.append(lst); pos += len(lst)
    chunks.append(np.full(8, -7, np.int32))
    pool = np.concatenate(chunks)
    return (pool, np.array(a_off, np.int64), np.array([len(a) for a, _ in pairs], np.int32),
            np.array(b_off, np.int64), np.array([len(b) for _, b in pairs], np.int32))


@pytest.fixture(scope="module")
def batch():
    pairs = make_pairs()
    pool, ao, al, bo, bl = pack(pairs)
    dev = torch.device("cuda:0")
    t = lambda x: torch.from_numpy(x).to(dev)
    rng = np.random.default_rng(5)
    bound, anc, anc2 = [], [], []
    for a, b in pairs:
        allv = np.concatenate([a, b, np.array([0], np.int32)])
        bound.append(int(rng.choice([0, int(np.median(allv)), int(allv.max()), int(allv.max()) + 1, 2**31 - 1])))
        inter = np.intersect1d(a, b)
        anc.append(int(inter[len(inter) // 2]) if len(inter) and rng.random() < 0.7 else int(a[0]) if len(a) else -1)
        anc2.append(int(inter[0]) if len(inter) and rng.random() < 0.5 else -1)
    return dict(pairs=pairs, pool=t(pool), ao=t(ao), al=t(al), bo=t(bo), bl=t(bl),
                bound=t(np.array(bound, np.int32)), anc=t(np.array(anc, np.int32)), anc2=t(np.array(anc2, np.int32)),
                h_bound=bound, h_anc=anc, h_anc2=anc2)


def run(batch, op, algo="bsearch", **kw):
    out = capi.intersect_batch(batch["pool"], batch["ao"], batch["al"], batch["bo"], batch["bl"], op=op, algo=algo,
                               bound=batch["bound"], anc=batch["anc"], anc2=batch["anc2"], **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("algo", ["auto", "bsearch", "merge", "hash", "gallop"])
def test_intersect_num_all_variants(batch, algo):
    got = run(batch, "intersect_num", algo)
    want = [oracle.intersection_num(a, b) for a, b in batch["pairs"]]
    assert got.tolist() == want


@pytest.mark.parametrize("algo", ["bsearch", "auto", "merge", "gallop", "hash"])
def test_counting_ops(batch, algo):
    """bounded / except / difference counts: the operator API (bsearch) and the streaming cores (merge-path,
    galloping, hash behind the TMA pipeline: lists truncated at the bound + O(log n) corrections)"""
    P, B, A, A2 = batch["pairs"], batch["h_bound"], batch["h_anc"], batch["h_anc2"]
    assert run(batch, "intersect_num_bound", algo).tolist() == [oracle.intersection_num(a, b, upper=u) for (a, b), u in zip(P, B)]
    assert run(batch, "intersect_num_bound_except", algo).tolist() == [
        oracle.intersection_num(a, b, upper=u, ancestors=(x,)) for (a, b), u, x in zip(P, B, A)]
    assert run(batch, "intersect_num_except2", algo).tolist() == [
        oracle.intersection_num(a, b, ancestors=(x, y)) for (a, b), x, y in zip(P, A, A2)]
    assert run(batch, "difference_num", algo).tolist() == [oracle.difference_num(a, b, x) for (a, b), x in zip(P, A)]
    assert run(batch, "difference_num_bound", algo).tolist() == [
        oracle.difference_num(a, b, x, upper=u) for (a, b), u, x in zip(P, B, A)]
    if algo in ("bsearch", "auto"):
        assert run(batch, "count_smaller", algo).tolist() == [oracle.bounded(a, u) for (a, _), u in zip(P, B)]


@pytest.mark.parametrize("op", ["intersect_set", "intersect_set_bound", "difference_set", "difference_set_bound"])
def test_materialising_ops(batch, op):
    P, B, A = batch["pairs"], batch["h_bound"], batch["h_anc"]
    cap = np.array([max(len(a), 1) for a, b in P], np.int64)
    off = np.zeros(len(P), np.int64); off[1:] = np.cumsum(cap)[:-1]
    out_pool = torch.full((int(cap.sum()),), -1, dtype=torch.int32, device="cuda:0")
    n = run(batch, op, out_pool=out_pool, out_off=torch.from_numpy(off).cuda())
    host = out_pool.cpu().numpy()
    for i, ((a, b), u, x) in enumerate(zip(P, B, A)):
        if op == "intersect_set":
            want = oracle.intersection_set(a, b)
        elif op == "intersect_set_bound":
            want = oracle.intersection_set(a, b, upper=u)
        elif op == "difference_set":
            want = oracle.difference_set(a, b, x)
        else:
            want = oracle.difference_set(a, b, x, upper=u)
        assert n[i] == len(want), (op, i)
        assert np.array_equal(host[off[i]: off[i] + n[i]], want), (op, i)


def test_streaming_checksum_at_scale():
    """Size-independent property at benchmark-like size: every variant returns the same per-pair
    counts on ~1M pairs (64M+ elements), and the total equals an independently computed checksum:
    list b_i is built as (a_i's even positions) ∪ (fresh odd values), so |a_i ∩ b_i| is known."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(1234)
    npairs, la = 1 << 18, 96
    base = torch.arange(npairs, device=dev, dtype=torch.int64)[:, None] * 0 \
        + torch.cumsum(torch.randint(1, 9, (npairs, la), device=dev, generator=g, dtype=torch.int64) * 2, 1)
    a = base                                                   # even, strictly increasing
    keep = torch.rand((npairs, la), device=dev, generator=g) < 0.4
    b = torch.where(keep, a, a + 1)                            # shared where keep, else odd (never in a)
    want = keep.sum(1)
    pool = torch.cat([a.reshape(-1), b.reshape(-1)]).to(torch.int32)
    ao = torch.arange(npairs, device=dev, dtype=torch.int64) * la
    bo = ao + npairs * la
    ln = torch.full((npairs,), la, dtype=torch.int32, device=dev)
    for algo in ("bsearch", "merge", "hash", "gallop"):
        got = capi.intersect_batch(pool, ao, ln, bo, ln, op="intersect_num", algo=algo)
        torch.cuda.synchronize()
        assert torch.equal(got, want), algo


def test_pipeline_all_size_classes_at_scale():
    """Every stage class of the TMA pipeline (one-, two- and four-warp groups, the over-long fallback,
    empty lists, unaligned starts) with enough pairs that every group wraps its stage ring many times:
    merge / gallop / auto must reproduce the operator-API counts pair by pair."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(99)
    npairs = 60000
    # lengths: mixture of tiny, small, medium, large and a few over-long lists; some empty
    cls = torch.randint(0, 6, (npairs, 2), device=dev, generator=g)
    hi = torch.tensor([1, 40, 300, 900, 2200, 5200], device=dev)[cls]
    ln = (torch.rand((npairs, 2), device=dev, generator=g) * hi.float()).long()
    gap = torch.randint(0, 4, (npairs, 2), device=dev, generator=g)              # unaligned list starts
    seg = (ln + gap).reshape(-1)
    off = torch.zeros(seg.numel() + 1, dtype=torch.int64, device=dev); torch.cumsum(seg, 0, out=off[1:])
    total = int(off[-1])
    starts = (off[:-1] + gap.reshape(-1))
    # strictly increasing values inside every list: global cumsum of gaps in {1,2,3}, rebased per list so
    # that a and b draw from the same range and intersect in about a third of their elements
    vals = torch.cumsum(torch.randint(1, 4, (total + 8,), device=dev, generator=g, dtype=torch.int64), 0)
    seg_id = torch.repeat_interleave(torch.arange(seg.numel(), device=dev), seg)
    base = vals[off[:-1]][seg_id]
    pool = torch.full((total + 8,), -7, dtype=torch.int32, device=dev)
    pool[:total] = (vals[:total] - base).to(torch.int32)
    ao, bo = starts[0::2].contiguous(), starts[1::2].contiguous()
    al, bl = ln[:, 0].to(torch.int32).contiguous(), ln[:, 1].to(torch.int32).contiguous()
    want = capi.intersect_batch(pool, ao, al, bo, bl, op="intersect_num", algo="bsearch")
    torch.cuda.synchronize()
    assert int(want.sum()) > 0
    for algo in ("merge", "gallop", "auto", "hash"):
        got = capi.intersect_batch(pool, ao, al, bo, bl, op="intersect_num", algo=algo)
        torch.cuda.synchronize()
        bad = torch.nonzero(got != want)
        assert bad.numel() == 0, (algo, bad[:5].tolist(), got[bad[:5, 0]].tolist(), want[bad[:5, 0]].tolist(),
                                  al[bad[:5, 0]].tolist(), bl[bad[:5, 0]].tolist())


@pytest.mark.parametrize("algo", ["auto", "merge", "gallop"])
def test_hub_pairs_take_the_cta_kernel(algo):
    """pairs beyond the largest TMA stage (4,608 staged elements) are intersected by one CTA each with a
    1,024-pivot shared-memory index (batch_list_cta_kernel)"""
    rng = np.random.default_rng(21)
    pairs = []
    for na, nb, hi in ((20000, 30000, 60000), (5000, 200000, 400000), (4700, 10, 9000), (100000, 100000, 120000), (6000, 0, 10)):
        pairs.append((np.unique(rng.integers(0, hi, na)).astype(np.int32), np.unique(rng.integers(0, hi, nb)).astype(np.int32)))
    pairs.append((pairs[0][0], pairs[0][0]))                        # identical hub rows
    pool, ao, al, bo, bl = pack(pairs)
    t = lambda x: torch.from_numpy(x).cuda()
    out = capi.intersect_batch(t(pool), t(ao), t(al), t(bo), t(bl), algo=algo)
    torch.cuda.synchronize()
    assert out.cpu().tolist() == [oracle.intersection_num(a, b) for a, b in pairs]
