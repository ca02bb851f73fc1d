"""The K-list stage of SearchByBoW drops every candidate at distance >= cut(TH_LOW, ratio) (csrc/matcher.cu bow_distance_cut).
This is sound iff no decision of the reference's per-row logic (src/ORBmatcher.cc:284-310, :1381-1400 for the KF-KF variant)

    accept  <=>  bestDist1 <= TH_LOW (or < TH_LOW)  and  (float)bestDist1 < ratio * (float)bestDist2

can depend on a candidate at or beyond the cut.  Enumerated here with the same float32 arithmetic, no GPU needed:
  (a) a candidate at distance >= cut is never an accepted best;
  (b) as second best it is indistinguishable from "no second candidate" (bestDist2 = 256, the reference's initial value)."""
import ctypes

import numpy as np
import pytest


def _accept(b1, b2, th_low, ratio, strict):
    ok = (b1 < th_low) if strict else (b1 <= th_low)
    return bool(ok and (np.float32(b1) < np.float32(ratio) * np.float32(b2)))


@pytest.mark.parametrize("th_low", [50, 100, 30, 0, 256])
@pytest.mark.parametrize("ratio", [0.6, 0.7, 0.75, 0.8, 0.9, 0.99, 1.0, 1.5, 0.3, 0.05])
def test_cut_is_sound(pkg, th_low, ratio):
    L = ctypes.CDLL(pkg.LIB_PATH)
    L.b2s_debug_bow_distance_cut.argtypes = [ctypes.c_int, ctypes.c_float]
    cut = L.b2s_debug_bow_distance_cut(th_low, ctypes.c_float(ratio))
    assert th_low + 2 <= cut <= 257 or cut == 257
    for strict in (False, True):
        for b1 in range(0, 257):
            if b1 >= cut:  # (a): never accepted, whatever the second best is
                assert not any(_accept(b1, b2, th_low, ratio, strict) for b2 in (b1, 256))
            for b2 in range(max(cut, b1), 257):  # (b): the decision with the true second best == the decision with 256
                assert _accept(b1, b2, th_low, ratio, strict) == _accept(b1, 256, th_low, ratio, strict), (b1, b2)


def test_no_cut_when_it_cannot_help(pkg):
    L = ctypes.CDLL(pkg.LIB_PATH)
    L.b2s_debug_bow_distance_cut.argtypes = [ctypes.c_int, ctypes.c_float]
    assert L.b2s_debug_bow_distance_cut(50, ctypes.c_float(0.7)) == 73
    assert L.b2s_debug_bow_distance_cut(100, ctypes.c_float(0.3)) == 257
    assert L.b2s_debug_bow_distance_cut(50, ctypes.c_float(0.0)) == 257
    assert L.b2s_debug_bow_distance_cut(50, ctypes.c_float(1.5)) == 52
