"""Pins ComputeStereoMatches, the feature grid and the SearchLocalPoints chain against the REFERENCE'S OWN Frame:
/root/reference/src/Frame.cc is compiled in place, unmodified, together with src/ORBextractor.cc and src/ORBmatcher.cc
(oracle/Makefile target `ref` -> oracle/_ref/libref_frame.so; MapPoint / KeyFrame / Converter / ORBVocabulary are
stand-ins).  A stereo Frame is built by the reference's constructor (src/Frame.cc:343-458); then

  * Frame::ComputeStereoMatches (:1026-1421): the oracle, given the keypoints / descriptors / pyramids the reference
    frame holds, must return the same mvuRight / mvDepth, float for float;
  * Frame::GetFeaturesInArea + AssignFeaturesToGrid (:461-491, :741-852): same indices in the same order as the oracle's
    grid (the candidate enumeration of every projection matcher);
  * Tracking::SearchLocalPoints (src/Tracking.cc:1166-1234): the reference's isInFrustum fills the track fields, its
    SearchByProjection(Frame&, points, th) matches; the oracle's orc_search_by_projection_map on those fields agrees."""
import ctypes
import os

import numpy as np
import pytest

from oracle_binding import FrameGeom, kp_dtype
from synth import synth_stereo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_frame.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")
vp, c_f, c_i = ctypes.c_void_p, ctypes.c_float, ctypes.c_int
FX, FY, CX, CY, BF = 718.856, 718.856, 607.1928, 185.2157, 386.1448
map_query_dtype = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("view_cos", "<f4"), ("level", "<i4"), ("in_view", "u1"),
                            ("has_obs", "u1"), ("pad", "u1", 2), ("desc", "u1", 32)])


def _kps(a):
    k = np.zeros(len(a), kp_dtype)
    for i, f in enumerate(("x", "y", "size", "angle", "response")):
        k[f] = a[:, i]
    k["octave"] = a[:, 5].astype(np.int32)
    k["class_id"] = a[:, 6].astype(np.int32)
    return k


class RefFrame:
    def __init__(self, imL, imR, nf, cx=CX, cy=CY):
        R = self.R = ctypes.CDLL(LIB)
        R.ref_frame_stereo.argtypes = [vp, vp, c_i, c_i, c_i, c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f, vp]
        R.ref_frame_get.argtypes = [vp] * 7
        R.ref_frame_get.restype = None
        R.ref_frame_level.argtypes = [c_i, c_i, vp, vp, vp]
        R.ref_frame_features_in_area.argtypes = [c_f, c_f, c_f, c_i, c_i, vp, c_i]
        R.ref_frame_search_local_points.argtypes = [vp, vp, vp, vp, vp, vp, c_i, c_f, vp, vp, vp, vp]
        imL, imR = np.ascontiguousarray(imL), np.ascontiguousarray(imR)
        h, w = imL.shape
        self.w, self.h = w, h
        nr = ctypes.c_int(0)
        self.N = R.ref_frame_stereo(imL.ctypes.data, imR.ctypes.data, w, h, nf, 1.2, 8, 20, 7, FX, FY, cx, cy, BF, 35.0, ctypes.byref(nr))
        self.NR = nr.value
        kl, kr = np.zeros((self.N, 7), np.float32), np.zeros((self.NR, 7), np.float32)
        self.descL, self.descR = np.zeros((self.N, 32), np.uint8), np.zeros((self.NR, 32), np.uint8)
        self.uright, self.depth = np.zeros(self.N, np.float32), np.zeros(self.N, np.float32)
        mb = ctypes.c_float(0)
        R.ref_frame_get(kl.ctypes.data, self.descL.ctypes.data, kr.ctypes.data, self.descR.ctypes.data, self.uright.ctypes.data,
                        self.depth.ctypes.data, ctypes.byref(mb))
        self.mb = mb.value
        self.kpsL, self.kpsR = _kps(kl), _kps(kr)

    def levels(self, right):
        out = []
        for l in range(8):
            w, h = ctypes.c_int(0), ctypes.c_int(0)
            self.R.ref_frame_level(right, l, None, ctypes.byref(w), ctypes.byref(h))
            a = np.zeros((h.value, w.value), np.uint8)
            self.R.ref_frame_level(right, l, a.ctypes.data, ctypes.byref(w), ctypes.byref(h))
            out.append(a)
        return out


def _oracle_stereo(oracle, F, mb):
    lv_l, lv_r = F.levels(0), F.levels(1)
    W = np.array([a.shape[1] for a in lv_l], np.int32)
    H = np.array([a.shape[0] for a in lv_l], np.int32)
    pl = (vp * 8)(*[a.ctypes.data for a in lv_l])
    pr = (vp * 8)(*[a.ctypes.data for a in lv_r])
    (sc, isc, _, _), _, _ = oracle.extractor(1000, 1.2, 8, 20, 7).tables()
    ur, dp = np.zeros(F.N, np.float32), np.zeros(F.N, np.float32)
    L = oracle.L
    L.orc_compute_stereo_matches.argtypes = [vp, vp, c_i, vp, vp, c_i, vp, vp, vp, vp, vp, vp, c_i, c_f, c_f, vp, vp]
    n = L.orc_compute_stereo_matches(F.kpsL.ctypes.data, F.descL.ctypes.data, F.N, F.kpsR.ctypes.data, F.descR.ctypes.data, F.NR,
                                     pl, pr, W.ctypes.data, H.ctypes.data, sc.ctypes.data, isc.ctypes.data, 8, BF, mb,
                                     ur.ctypes.data, dp.ctypes.data)
    return n, ur, dp


@pytest.mark.parametrize("w,h,nf,seed", [(640, 480, 1000, 9), (1241, 376, 2000, 3), (752, 480, 1200, 21)])
def test_compute_stereo_matches_equals_reference_frame(oracle, w, h, nf, seed):
    imL, imR = synth_stereo(w, h, seed)
    F = RefFrame(imL, imR, nf)
    assert F.N > nf * 0.8 and F.NR > nf * 0.8
    # this fork assigns mb AFTER ComputeStereoMatches ran inside the constructor (src/Frame.cc:403-449): the call sees 0
    assert abs(F.mb - BF / FX) < 1e-6
    n, ur, dp = _oracle_stereo(oracle, F, 0.0)
    assert n == int((F.uright >= 0).sum()) and n > 100
    assert np.array_equal(ur, F.uright)
    assert np.array_equal(dp, F.depth)


def test_features_in_area_equals_reference_frame(oracle):
    imL, imR = synth_stereo(1241, 376, 5)
    F = RefFrame(imL, imR, 2000)
    (sc, _, _, _), _, _ = oracle.extractor(2000, 1.2, 8, 20, 7).tables()
    g = FrameGeom(0.0, 0.0, float(F.w), float(F.h), BF, sc.ctypes.data, 8)
    kx, ky = np.ascontiguousarray(F.kpsL["x"]), np.ascontiguousarray(F.kpsL["y"])
    ko = np.ascontiguousarray(F.kpsL["octave"])
    L = oracle.L
    L.orc_features_in_area.argtypes = [vp, vp, vp, c_i, vp, c_f, c_f, c_f, c_i, c_i, vp, c_i]
    rng = np.random.RandomState(3)
    a, b = np.zeros(4096, np.int32), np.zeros(4096, np.int32)
    total = 0
    for t in range(600):
        if t % 3 == 0:  # around a feature, else anywhere (also outside the image)
            j = rng.randint(0, F.N)
            x, y = float(kx[j]) + rng.uniform(-3, 3), float(ky[j]) + rng.uniform(-3, 3)
        else:
            x, y = rng.uniform(-60, F.w + 60), rng.uniform(-60, F.h + 60)
        r = float(rng.choice([2.5, 7.0, 15.0, 40.0, 120.0]))
        mn, mx = [(-1, -1), (0, 3), (2, -1), (3, 4), (0, 0), (1, 7)][rng.randint(0, 6)]
        na = F.R.ref_frame_features_in_area(x, y, r, mn, mx, a.ctypes.data, 4096)
        nb = L.orc_features_in_area(kx.ctypes.data, ky.ctypes.data, ko.ctypes.data, F.N, ctypes.byref(g), x, y, r, mn, mx,
                                    b.ctypes.data, 4096)
        assert na == nb and np.array_equal(a[:na], b[:nb])
        total += na
    assert total > 5000


@pytest.mark.parametrize("seed,th", [(5, 1.0), (6, 3.0)])
def test_search_local_points_chain_equals_reference(oracle, checker, seed, th):
    imL, imR = synth_stereo(1241, 376, seed)
    F = RefFrame(imL, imR, 2000)
    rng = np.random.RandomState(seed)
    (sc, _, _, _), _, _ = oracle.extractor(2000, 1.2, 8, 20, 7).tables()
    # a pose and map points that project near the frame's features
    ang = 0.02
    Rm = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.2, -0.1, 0.4])
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = Rm, t
    T = np.ascontiguousarray(T, np.float32)
    n = 2500
    src = rng.randint(0, F.N, size=n)
    z = rng.uniform(3, 50, n)
    u = F.kpsL["x"][src] + rng.normal(0, 1.5, n)
    v = F.kpsL["y"][src] + rng.normal(0, 1.0, n)
    Xc = np.stack([(u - CX) * z / FX, (v - CY) * z / FY, z], 1)
    Xw = (Rm.T @ (Xc - t).T).T
    Ow = -Rm.T @ t
    PO = Xw - Ow
    dist = np.linalg.norm(PO, axis=1)
    nrm = PO / dist[:, None] + rng.normal(0, 0.35, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    nrm[rng.randint(0, 100, size=n) < 5] *= -1
    lvl = np.clip(F.kpsL["octave"][src] + rng.randint(0, 2, size=n), 0, 7)
    maxd = (dist * sc[lvl] * rng.uniform(0.9, 1.0, n)).astype(np.float32)
    maxd[rng.randint(0, 100, size=n) < 5] *= 0.3
    mind = (maxd / sc[7]).astype(np.float32)
    behind = rng.randint(0, 100, size=n) < 3
    Xw[behind] = Ow + (Ow - Xw[behind])
    desc = F.descL[src].copy()
    for i in range(n):
        for b in rng.choice(256, size=int(rng.randint(0, 40)), replace=False):
            desc[i, b >> 3] ^= np.uint8(1 << (b & 7))
    pos, nrm = np.ascontiguousarray(Xw, np.float32), np.ascontiguousarray(nrm, np.float32)
    match = np.full(F.N, -1, np.int32)
    track = np.zeros((n, 4), np.float32)
    level = np.zeros(n, np.int32)
    inview = np.zeros(n, np.uint8)
    nr = F.R.ref_frame_search_local_points(T.ctypes.data, pos.ctypes.data, nrm.ctypes.data, mind.ctypes.data, maxd.ctypes.data,
                                           desc.ctypes.data, n, th, match.ctypes.data, track.ctypes.data, level.ctypes.data,
                                           inview.ctypes.data)
    assert 0.5 * n < inview.sum() < n
    q = np.zeros(n, map_query_dtype)
    q["u"], q["v"], q["ur"], q["view_cos"] = track[:, 0], track[:, 1], track[:, 2], track[:, 3]
    q["level"], q["in_view"], q["has_obs"], q["desc"] = level, inview, 1, desc
    geom = dict(mnMinX=np.float32(0), mnMinY=np.float32(0), mnMaxX=np.float32(F.w), mnMaxY=np.float32(F.h), bf=np.float32(BF),
                scale_factors=sc)
    no, mo = checker.search_by_projection_map(q, np.ascontiguousarray(F.kpsL["x"]), np.ascontiguousarray(F.kpsL["y"]),
                                             np.ascontiguousarray(F.kpsL["octave"]), F.uright, np.zeros(F.N, np.uint8), F.descL,
                                             geom, th=th, th_high=100, nnratio=0.8)
    assert nr == no and nr > 300
    assert np.array_equal(match, mo)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,nf,seed", [(640, 480, 1000, 9), (1241, 376, 2000, 3)])
def test_cuda_stereo_matches_equal_reference_frame(pkg, w, h, nf, seed):
    """The CUDA ComputeStereoMatches kernel against the reference Frame directly: the device extracts the same stereo pair
    (its resident pyramids are the reference's, bit for bit), is handed the REFERENCE frame's keypoints and descriptors, and
    must return the reference frame's mvuRight / mvDepth."""
    import torch
    imL, imR = synth_stereo(w, h, seed)
    F = RefFrame(imL, imR, nf)
    ex = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2)
    ex.extract_batch([imL, imR])
    cap = ex.cap
    assert F.N <= cap and F.NR <= cap
    kps = np.zeros((2, cap), kp_dtype)
    desc = np.zeros((2, cap, 32), np.uint8)
    kps[0, :F.N], kps[1, :F.NR] = F.kpsL, F.kpsR
    desc[0, :F.N], desc[1, :F.NR] = F.descL, F.descR
    d_kps = torch.from_numpy(kps.view(np.uint8).reshape(2, cap, 28)).cuda()
    d_desc = torch.from_numpy(desc).cuda()
    d_cnt = torch.tensor([F.N, F.NR], dtype=torch.int32).cuda()
    ur = torch.full((1, cap), -2.0, dtype=torch.float32, device="cuda")
    dp = torch.full((1, cap), -2.0, dtype=torch.float32, device="cuda")
    nm = torch.zeros(1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ex.stereo_match_device(0, 1, 1, d_kps.data_ptr(), d_desc.data_ptr(), d_cnt.data_ptr(), cap, BF, 0.0, ur.data_ptr(),
                           dp.data_ptr(), nm.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert int(nm.item()) == int((F.uright >= 0).sum()) and int(nm.item()) > 100
    assert np.array_equal(ur.cpu().numpy()[0, :F.N], F.uright)
    assert np.array_equal(dp.cpu().numpy()[0, :F.N], F.depth)
