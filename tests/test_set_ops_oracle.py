"""Set-operator semantics of the oracle against brute-force numpy set algebra (the reference never
unit-tests its operators)."""
import numpy as np
import pytest

import oracle


def cases():
    rng = np.random.default_rng(7)
    out = [(np.array([], np.int32), np.array([], np.int32)),
           (np.array([5], np.int32), np.array([], np.int32)),
           (np.array([], np.int32), np.array([1, 2, 3], np.int32)),
           (np.array([3], np.int32), np.array([3], np.int32)),
           (np.arange(0, 50, 2, dtype=np.int32), np.arange(1, 51, 2, dtype=np.int32)),     # disjoint
           (np.arange(40, dtype=np.int32), np.arange(40, dtype=np.int32))]                 # identical
    for na, nb, hi in [(10, 10, 30), (100, 7, 300), (3, 1000, 3000), (500, 500, 1200), (70, 2000, 2500)]:
        a = np.unique(rng.integers(0, hi, na)).astype(np.int32)
        b = np.unique(rng.integers(0, hi, nb)).astype(np.int32)
        out.append((a, b))
    return out


@pytest.mark.parametrize("a,b", cases())
def test_operators_match_set_algebra(a, b):
    inter = np.intersect1d(a, b)
    assert oracle.intersection_num(a, b) == len(inter)
    assert np.array_equal(oracle.intersection_set(a, b), inter)
    bounds = sorted(set([0, 1, 10**9] + ([int(a[len(a) // 2])] if len(a) else []) + ([int(b[-1])] if len(b) else [])))
    for up in bounds:
        assert oracle.intersection_num(a, b, upper=up) == int((inter < up).sum())
        assert np.array_equal(oracle.intersection_set(a, b, upper=up), inter[inter < up])
        assert oracle.bounded(a, up) == int((a < up).sum())
        for vid in ([-1] + ([int(a[0])] if len(a) else [])):
            diff = np.setdiff1d(a, b)
            diff = diff[diff != vid]
            assert oracle.difference_num(a, b, vid) == len(diff)
            assert np.array_equal(oracle.difference_set(a, b, vid), diff)
            assert oracle.difference_num(a, b, vid, upper=up) == int((diff < up).sum())
            assert np.array_equal(oracle.difference_set(a, b, vid, upper=up), diff[diff < up])
    if len(inter):
        x, y = int(inter[0]), int(inter[-1])
        assert oracle.intersection_num(a, b, ancestors=(x,)) == len(inter) - 1
        assert oracle.intersection_num(a, b, ancestors=(x, y)) == len(inter) - (1 if x == y else 2)
        assert oracle.intersection_num(a, b, upper=y, ancestors=(x,)) == int((inter < y).sum()) - (1 if x < y else 0)
        assert np.array_equal(oracle.intersection_set(a, b, ancestor=x), inter[inter != x])
