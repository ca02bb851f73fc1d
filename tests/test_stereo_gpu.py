"""GPU parity of Frame::ComputeStereoMatches (src/Frame.cc:1026-1421) against the CPU oracle:
mvuRight / mvDepth bit-exact (float), same surviving match count."""
import numpy as np
import pytest

from synth import synth_stereo

pytestmark = pytest.mark.gpu

BF = 386.1448
MB_REAL = 386.1448 / 718.856


@pytest.mark.parametrize("w,h,nfeat,seed", [(1241, 376, 2000, 5), (640, 480, 1000, 9)])
@pytest.mark.parametrize("mb", [0.0, MB_REAL, BF / 30.0])
def test_compute_stereo_matches(pkg, oracle, w, h, nfeat, seed, mb):
    left, right = synth_stereo(w, h, seed)
    ex = pkg.ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2)
    (kL, dL), (kR, dR) = ex.extract_batch([left, right])
    ur, dp, nm = ex.stereo_match(0, 1, 1, BF, mb)
    oL, oR = oracle.extractor(nfeat, 1.2, 8, 20, 7), oracle.extractor(nfeat, 1.2, 8, 20, 7)
    okL, odL = oL(left)
    okR, odR = oR(right)
    assert np.array_equal(okL, kL) and np.array_equal(odR, dR)  # same features on both sides (extractor parity)
    n, our, odp = oracle.compute_stereo_matches(oL, oR, okL, odL, okR, odR, BF, mb)
    assert n == int(nm[0])
    assert np.array_equal(ur[0, :len(kL)], our)
    assert np.array_equal(dp[0, :len(kL)], odp)
    assert n > 50  # the synthetic pair really matches (fewer with the tight disparity bound)
    assert (ur[0, len(kL):] == -1).all()


def test_stereo_match_batch_of_pairs(pkg, oracle):
    """Several pairs in one call (left images first, then right images), device-resident records."""
    w, h, nfeat, F = 640, 480, 1000, 3
    pairs = [synth_stereo(w, h, 20 + i) for i in range(F)]
    ex = pkg.ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2 * F)
    res = ex.extract_batch([p[0] for p in pairs] + [p[1] for p in pairs])
    ur, dp, nm = ex.stereo_match(0, F, F, BF, 0.0)
    for i in range(F):
        oL, oR = oracle.extractor(nfeat, 1.2, 8, 20, 7), oracle.extractor(nfeat, 1.2, 8, 20, 7)
        okL, odL = oL(pairs[i][0])
        okR, odR = oR(pairs[i][1])
        n, our, odp = oracle.compute_stereo_matches(oL, oR, okL, odL, okR, odR, BF, 0.0)
        assert n == int(nm[i]) and np.array_equal(ur[i, :len(okL)], our) and np.array_equal(dp[i, :len(okL)], odp)


def test_stereo_match_requires_extraction(pkg):
    ex = pkg.ORBextractor(500, 1.2, 8, 20, 7, max_width=320, max_height=240, max_batch=2)
    with pytest.raises(pkg.B200SlamError):
        ex.stereo_match(0, 1, 1, BF)
