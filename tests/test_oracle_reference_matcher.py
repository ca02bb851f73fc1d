"""Pins the oracle's matchers against the REFERENCE'S OWN SOURCE: /root/reference/src/ORBmatcher.cc is compiled in place,
unmodified (oracle/Makefile target `ref` -> oracle/_ref/libref_matcher.so), against oracle/refshim — the cv stand-in
plus data-holder versions of Frame / KeyFrame / MapPoint (refshim/slam_stubs.h) — and driven through
oracle/ref_matcher_glue.cpp on flattened inputs.  Every matcher the CUDA path implements must give the same matches as
the oracle function the GPU parity tests use:

  SearchByBoW(KeyFrame*, Frame&)                   src/ORBmatcher.cc:230    <-> orc_search_by_bow (accept <= TH_LOW)
  SearchByBoW(KeyFrame*, KeyFrame*)                :656                     <-> orc_search_by_bow (accept <  TH_LOW)
  SearchByProjection(Frame&, vpMapPoints, th)      :70                      <-> orc_search_by_projection_map
  SearchByProjection(Current, Last, th, bMono)     :1569                    <-> orc_search_by_projection_last
  SearchByProjection(Current, KeyFrame*, found,..) :1731 (relocalisation)   <-> orc_search_by_projection_last, mode 0, no stereo gate
  SearchForInitialization                          :515                     <-> orc_search_for_initialization
  SearchForTriangulation                           :810                     <-> orc_search_for_triangulation
  Fuse(KeyFrame*, vpMapPoints, th)                 :1020                    <-> orc_search_windows (CHI2)
  Fuse(KeyFrame*, Scw, vpPoints, th, vpReplace)    :1179                    <-> orc_search_windows
  SearchByProjection(KeyFrame*, Scw, ...)          :388                     <-> orc_search_windows (GREEDY)
  SearchBySim3                                     :1314                    <-> orc_search_windows x2 + mutual check

Where the reference projects the map points itself, the glue exports the post-projection queries (same expressions on
the same cv stand-in) and the oracle consumes those — so the comparison covers candidate enumeration, gates, distances,
tie-breaking, greedy state and the rotation-histogram cull, bit for bit.

Every case runs twice (conftest.py `checker`): oracle vs reference source in the CPU suite, and — marked gpu — the CUDA
kernels vs the reference source directly."""
import ctypes
import os

import numpy as np
import pytest

from oracle_binding import proj_query_dtype
from synth import synth_descriptors, synth_projection, synth_projection_map, synth_triangulation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_matcher.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")
vp = ctypes.c_void_p
c_f, c_i = ctypes.c_float, ctypes.c_int32

win_query_dtype = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"),
                            ("max_level", "<i4"), ("valid", "u1"), ("pad", "u1", 3), ("desc", "u1", 32)])
map_query_dtype = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("view_cos", "<f4"), ("level", "<i4"), ("in_view", "u1"),
                            ("has_obs", "u1"), ("pad", "u1", 2), ("desc", "u1", 32)])

FX, FY, CX, CY, BF = 718.856, 718.856, 607.1928, 185.2157, 386.1448
W, H = 1241, 376


class Cam(ctypes.Structure):
    _fields_ = [("fx", c_f), ("fy", c_f), ("cx", c_f), ("cy", c_f), ("bf", c_f), ("b", c_f), ("minX", c_f), ("minY", c_f),
                ("maxX", c_f), ("maxY", c_f), ("nlevels", c_i), ("scaleFactor", c_f)]


class Feats(ctypes.Structure):
    _fields_ = [("n", c_i), ("x", vp), ("y", vp), ("angle", vp), ("octave", vp), ("uright", vp), ("desc", vp), ("node", vp),
                ("mp", vp), ("outlier", vp), ("Tcw", vp)]


class Points(ctypes.Structure):
    _fields_ = [("n", c_i), ("pos", vp), ("normal", vp), ("desc", vp), ("bad", vp), ("nobs", vp), ("minDist", vp),
                ("maxDist", vp), ("trackX", vp), ("trackY", vp), ("trackXR", vp), ("viewCos", vp), ("trackLevel", vp),
                ("inView", vp)]


def cam():
    return Cam(FX, FY, CX, CY, BF, BF / FX, 0.0, 0.0, float(W), float(H), 8, 1.2)


def scale_tables():
    """mvScaleFactors / mvLevelSigma2 as src/ORBextractor.cc:468-491 builds them (running float product)."""
    s = [np.float32(1.0)]
    for _ in range(7):
        s.append(np.float32(s[-1] * np.float32(1.2)))
    s = np.array(s, np.float32)
    return s, (s * s).astype(np.float32)


class Keep:
    """Builds the ctypes structs and keeps the numpy arrays alive."""

    def __init__(self):
        self.hold = []

    def arr(self, a, dt):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dt)
        self.hold.append(a)
        return a.ctypes.data

    def feats(self, x, y, angle, octave, uright, desc, node=None, mp=None, outlier=None, Tcw=None):
        n = len(x)
        return Feats(n, self.arr(x, np.float32), self.arr(y, np.float32), self.arr(angle, np.float32), self.arr(octave, np.int32),
                     self.arr(uright, np.float32), self.arr(desc, np.uint8), self.arr(node, np.int32), self.arr(mp, np.int32),
                     self.arr(outlier, np.uint8), self.arr(Tcw, np.float32))

    def points(self, pos, desc, normal=None, bad=None, nobs=None, minDist=None, maxDist=None, track=None):
        n = len(pos)
        t = track or {}
        return Points(n, self.arr(pos, np.float32), self.arr(normal, np.float32), self.arr(desc, np.uint8), self.arr(bad, np.uint8),
                      self.arr(nobs, np.int32), self.arr(minDist, np.float32), self.arr(maxDist, np.float32),
                      self.arr(t.get("x"), np.float32), self.arr(t.get("y"), np.float32), self.arr(t.get("xr"), np.float32),
                      self.arr(t.get("cos"), np.float32), self.arr(t.get("level"), np.int32), self.arr(t.get("in_view"), np.uint8))


_SIGS = {
    "ref_search_by_bow_kf_f": [vp, vp, vp, vp, c_f, c_i, vp],
    "ref_search_by_bow_kf_kf": [vp, vp, vp, vp, c_f, c_i, vp],
    "ref_search_by_projection_map": [vp, vp, vp, vp, c_i, c_f, c_f, vp],
    "ref_search_by_projection_last": [vp, vp, vp, vp, c_f, c_i, c_i, vp, vp, vp],
    "ref_search_for_triangulation": [vp, vp, vp, vp, vp, c_i, c_i, vp, vp],
    "ref_search_for_initialization": [vp, vp, vp, vp, c_i, c_f, c_i, vp],
    "ref_search_by_projection_reloc": [vp, vp, vp, vp, vp, c_f, c_i, c_i, vp, vp],
    "ref_fuse": [vp, vp, vp, vp, c_i, vp, c_f, vp, vp],
    "ref_search_by_projection_scw": [vp, vp, vp, vp, c_i, vp, vp, c_i, vp, vp],
    "ref_search_by_sim3": [vp, vp, vp, vp, vp, c_f, vp, vp, c_f, vp, vp, vp],
    "ref_descriptor_distance": [vp, vp],
}
# result arrays of every entry point: (argument index, numpy dtype, element count from the call's arguments); `inout`
# arrays are restored before the second library runs
_N = lambda i: (lambda a: a[i]._obj.n)
_OUTS = {
    "ref_search_by_bow_kf_f": [(6, np.int32, _N(3))],
    "ref_search_by_bow_kf_kf": [(6, np.int32, _N(1))],
    "ref_search_by_projection_map": [(7, np.int32, _N(1))],
    "ref_search_by_projection_last": [(7, np.int32, _N(1)), (9, np.int32, lambda a: 1)],
    "ref_search_by_projection_reloc": [(8, np.int32, _N(1))],
    "ref_search_for_initialization": [(7, np.int32, _N(1)), (3, np.float32, lambda a: 2 * a[1]._obj.n)],
    "ref_search_for_triangulation": [(7, np.int32, _N(1)), (8, np.float32, lambda a: 2)],
    "ref_fuse": [(7, np.int32, lambda a: a[4])],
    "ref_search_by_projection_scw": [(8, np.int32, lambda a: a[4])],
    "ref_search_by_sim3": [(9, np.int32, _N(1))],
}
ADAPTER_LIB = os.path.join(ROOT, "oracle", "_ref", "libadapter_matcher.so")


def _view(ptr, dt, n):
    addr = ctypes.addressof(ptr._obj) if hasattr(ptr, "_obj") else int(ptr)  # byref(x) or a raw address
    return np.ctypeslib.as_array((ctypes.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(addr)).view(dt)


class _RefAndAdapter:
    """The reference's ORBmatcher (libref_matcher.so) and — when a CUDA device is present — the product's drop-in
    ORBmatcher with the same C++ signatures (host/adapters/ORBmatcher_b200.cc in libadapter_matcher.so), driven by the
    same glue on the same stand-in objects: every ref_* call is repeated as adp_* and must return the same count and the
    same result arrays."""

    def __init__(self, R, A):
        self.R, self.A = R, A
        self.adapter_calls = 0

    def __getattr__(self, name):
        fr = getattr(self.R, name)
        if self.A is None or name not in _OUTS:
            return fr
        fa = getattr(self.A, "adp_" + name[4:])

        def both(*args):
            outs = [(_view(args[i], dt, cnt(args)), i) for i, dt, cnt in _OUTS[name]]
            before = [o.copy() for o, _ in outs]
            nr = fr(*args)
            want = [o.copy() for o, _ in outs]
            for (o, _), b0 in zip(outs, before):
                o[...] = b0
            na = fa(*args)
            self.adapter_calls += 1
            assert na == nr, "%s: the adapter returns %d, the reference %d" % (name, na, nr)
            for (o, i), w in zip(outs, want):
                assert np.array_equal(o, w), "%s: result array (argument %d) differs between adapter and reference" % (name, i)
            return nr
        return both


@pytest.fixture(scope="module")
def ref(request):
    R = ctypes.CDLL(LIB)
    A = None
    if os.path.exists(ADAPTER_LIB):
        try:
            pkg = request.getfixturevalue("pkg")
            if pkg.device_count() > 0:
                A = ctypes.CDLL(ADAPTER_LIB)
        except Exception:
            A = None
    for name, sig in _SIGS.items():
        getattr(R, name).restype = ctypes.c_int
        getattr(R, name).argtypes = sig
        if A is not None:
            getattr(A, "adp_" + name[4:]).restype = ctypes.c_int
            getattr(A, "adp_" + name[4:]).argtypes = sig
    return _RefAndAdapter(R, A)


def B(x):
    return ctypes.byref(x)


def test_descriptor_distance(ref, oracle):
    rng = np.random.RandomState(1)
    a = rng.randint(0, 256, size=(200, 32)).astype(np.uint8)
    b = rng.randint(0, 256, size=(200, 32)).astype(np.uint8)
    b[:20] = a[:20]
    for i in range(200):
        assert ref.ref_descriptor_distance(a[i].ctypes.data, b[i].ctypes.data) == oracle.descriptor_distance(a[i], b[i])


# ---------------------------------------------------------------- SearchByBoW
def _bow_inputs(n, seed, match_frac):
    A, nodeA, validA, angA, Bd, nodeB, angB = synth_descriptors(n, seed=seed, match_frac=match_frac)
    rng = np.random.RandomState(seed + 1)
    xy = rng.uniform(0, 300, size=(4, n)).astype(np.float32)
    octv = rng.randint(0, 8, size=n).astype(np.int32)
    ur = np.full(n, -1, np.float32)
    return A, nodeA, validA, angA, Bd, nodeB, angB, xy, octv, ur, rng


@pytest.mark.parametrize("n,seed,match_frac,check_ori", [(1500, 11, 0.7, 1), (1500, 12, 0.3, 1), (600, 13, 0.9, 0)])
def test_search_by_bow_keyframe_frame(ref, checker, n, seed, match_frac, check_ori):
    A, nodeA, validA, angA, Bd, nodeB, angB, xy, octv, ur, rng = _bow_inputs(n, seed, match_frac)
    k = Keep()
    # invalid keyframe features: half without a map point, half with a bad one
    mpA = np.arange(n, dtype=np.int32)
    inv = np.nonzero(validA == 0)[0]
    mpA[inv[::2]] = -1
    bad = np.zeros(n, np.uint8)
    bad[inv[1::2]] = 1
    pts = k.points(np.zeros((n, 3), np.float32), A, bad=bad)
    kf = k.feats(xy[0], xy[1], angA, octv, ur, A, node=nodeA, mp=mpA)
    fr = k.feats(xy[2], xy[3], angB, octv, ur, Bd, node=nodeB)
    m = np.full(n, -1, np.int32)
    c = cam()
    nr = ref.ref_search_by_bow_kf_f(B(c), B(kf), B(pts), B(fr), 0.7, check_ori, m.ctypes.data)
    no, mo = checker.search_by_bow(A, nodeA, validA, angA, Bd, nodeB, angB, th_low=50, nnratio=0.7, strict_lt=False,
                                  check_ori=bool(check_ori))
    assert nr == no and nr > n * match_frac * 0.3
    assert np.array_equal(m, mo)


@pytest.mark.parametrize("n,seed,match_frac,check_ori", [(1500, 21, 0.7, 1), (800, 22, 0.5, 0)])
def test_search_by_bow_keyframe_keyframe(ref, checker, n, seed, match_frac, check_ori):
    A, nodeA, validA, angA, Bd, nodeB, angB, xy, octv, ur, rng = _bow_inputs(n, seed, match_frac)
    validB = (rng.randint(0, 100, size=n) < 90).astype(np.uint8)
    k = Keep()
    mp1 = np.where(validA != 0, np.arange(n), -1).astype(np.int32)
    mp2 = (np.arange(n) + n).astype(np.int32)
    inv = np.nonzero(validB == 0)[0]
    mp2[inv[::2]] = -1
    bad = np.zeros(2 * n, np.uint8)
    bad[n + inv[1::2]] = 1
    pts = k.points(np.zeros((2 * n, 3), np.float32), np.concatenate([A, Bd]), bad=bad)
    kf1 = k.feats(xy[0], xy[1], angA, octv, ur, A, node=nodeA, mp=mp1)
    kf2 = k.feats(xy[2], xy[3], angB, octv, ur, Bd, node=nodeB, mp=mp2)
    m12 = np.full(n, -1, np.int32)
    c = cam()
    nr = ref.ref_search_by_bow_kf_kf(B(c), B(kf1), B(kf2), B(pts), 0.75, check_ori, m12.ctypes.data)
    no, mB = checker.search_by_bow(A, nodeA, validA, angA, Bd, nodeB, angB, validB=validB, th_low=50, nnratio=0.75,
                                  strict_lt=True, check_ori=bool(check_ori))
    assert nr == no and nr > 50
    # the oracle reports per kf2 feature, the reference per kf1 feature
    inv12 = np.full(n, -1, np.int32)
    j = np.nonzero(mB >= 0)[0]
    inv12[mB[j]] = j
    assert np.array_equal(m12, inv12)


# ---------------------------------------------------------------- SearchByProjection(Frame&, local map points)
@pytest.mark.parametrize("seed,cluster,th", [(31, False, 1.0), (32, True, 3.0), (33, False, 5.0)])
def test_search_by_projection_map(ref, checker, seed, cluster, th):
    d = synth_projection_map(nf=1500, nq=1800, seed=seed, cluster=cluster)
    sf, _ = scale_tables()
    q, nf, nq = d["q"], len(d["kpx"]), len(d["q"])
    rng = np.random.RandomState(seed)
    # not-in-view queries: half flagged by mbTrackInView, half by isBad()
    out = np.nonzero(q["in_view"] == 0)[0]
    in_view = np.ones(nq, np.uint8)
    bad = np.zeros(nq, np.uint8)
    in_view[out[::2]] = 0
    bad[out[1::2]] = 1
    occ = np.nonzero(d["occupied"])[0]
    free_with_mp = np.nonzero(d["occupied"] == 0)[0][::9]  # features that hold a point nobody observes: not occupied
    nocc, nfree = len(occ), len(free_with_mp)
    npts = nq + nocc + nfree
    mp = np.full(nf, -1, np.int32)
    mp[occ] = nq + np.arange(nocc)
    mp[free_with_mp] = nq + nocc + np.arange(nfree)
    nobs = np.concatenate([q["has_obs"].astype(np.int32), np.ones(nocc, np.int32), np.zeros(nfree, np.int32)])

    def ext(a, fill=0):
        return np.concatenate([a, np.full(npts - nq, fill, a.dtype)])
    k = Keep()
    pts = k.points(np.zeros((npts, 3), np.float32), np.concatenate([q["desc"], np.zeros((npts - nq, 32), np.uint8)]),
                   bad=ext(bad), nobs=nobs,
                   track=dict(x=ext(q["u"]), y=ext(q["v"]), xr=ext(q["ur"]), cos=ext(q["view_cos"]), level=ext(q["level"]),
                              in_view=ext(in_view)))
    fr = k.feats(d["kpx"], d["kpy"], np.zeros(nf, np.float32), d["octave"], d["uright"], d["desc"], mp=mp)
    qi = np.arange(nq, dtype=np.int32)
    m = np.full(nf, -1, np.int32)
    c = cam()
    nr = ref.ref_search_by_projection_map(B(c), B(fr), B(pts), qi.ctypes.data, nq, th, 0.8, m.ctypes.data)
    g = dict(d["geom"])
    g["scale_factors"] = sf
    no, mo = checker.search_by_projection_map(q, d["kpx"], d["kpy"], d["octave"], d["uright"], d["occupied"], d["desc"], g, th=th,
                                             th_high=100, nnratio=0.8)
    assert nr == no and nr > 200
    assert np.array_equal(m, mo)


# ---------------------------------------------------------------- helpers: a small 3-D world
def _rot(rx, ry, rz):
    cx_, sx = np.cos(rx), np.sin(rx)
    cy_, sy = np.cos(ry), np.sin(ry)
    cz, sz = np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, cx_, -sx], [0, sx, cx_]])
    Ry = np.array([[cy_, 0, sy], [0, 1, 0], [-sy, 0, cy_]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _pose(R, t):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return T.astype(np.float32)


def _flip(rng, d, kmax):
    d = d.copy()
    for b in rng.choice(256, size=int(rng.randint(0, kmax)), replace=False):
        d[b >> 3] ^= np.uint8(1 << (b & 7))
    return d


def _features(rng, nf, cluster=False):
    if cluster:
        x = (300 + rng.randint(0, 1200, size=nf) / 10.0).astype(np.float32)
        y = (150 + rng.randint(0, 600, size=nf) / 10.0).astype(np.float32)
        base = rng.randint(0, 256, size=(4, 32)).astype(np.uint8)
        desc = np.stack([_flip(rng, base[rng.randint(0, 4)], 12) for _ in range(nf)])
    else:
        x = (rng.randint(0, W * 10, size=nf) / 10.0).astype(np.float32)
        y = (rng.randint(0, H * 10, size=nf) / 10.0).astype(np.float32)
        desc = rng.randint(0, 256, size=(nf, 32)).astype(np.uint8)
    octv = rng.randint(0, 8, size=nf).astype(np.int32)
    ang = (rng.randint(0, 360000, size=nf) / 1000.0).astype(np.float32)
    depth = rng.uniform(4, 60, size=nf)
    ur = (x - BF / depth).astype(np.float32)
    ur[rng.randint(0, 100, size=nf) < 20] = -1.0
    return dict(x=x, y=y, octave=octv, angle=ang, desc=desc, uright=ur, depth=depth)


def _points_seen_from(rng, f, Tcw, n, sf, px_noise=1.5, flips=30, S=None, src=None):
    """n map points that project near features of `f` under pose Tcw (or under the similarity S): position, normal,
    descriptor, distance-invariance range consistent with the source feature's level."""
    nf = len(f["x"])
    src = rng.randint(0, nf, size=n) if src is None else src
    u = f["x"][src].astype(np.float64) + rng.normal(0, px_noise, n)
    v = f["y"][src].astype(np.float64) + rng.normal(0, px_noise * 0.5, n)
    z = f["depth"][src] * rng.uniform(0.97, 1.03, n)
    Xc = np.stack([(u - CX) * z / FX, (v - CY) * z / FY, z], 1)
    if S is None:
        R, t = Tcw[:3, :3].astype(np.float64), Tcw[:3, 3].astype(np.float64)
        Xw = (R.T @ (Xc - t).T).T
        Ow = -R.T @ t
    else:
        sR, t = S[:3, :3].astype(np.float64), S[:3, 3].astype(np.float64)
        s = np.sqrt((sR[0] ** 2).sum())
        R, t = sR / s, t / s
        Xw = (R.T @ (Xc - t).T).T  # the reference projects with the scale-free part
        Ow = -R.T @ t
    PO = Xw - Ow
    dist = np.linalg.norm(PO, axis=1)
    nrm = PO / dist[:, None] + rng.normal(0, 0.25, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    far = rng.randint(0, 100, size=n) < 6  # viewed from the side: fails the 60-degree test
    nrm[far] = -nrm[far]
    lvl = np.clip(f["octave"][src] + rng.randint(0, 2, size=n), 0, 7)
    maxd = dist * sf[lvl] * rng.uniform(0.9, 1.0, n)
    mind = maxd / sf[7]
    rng_out = rng.randint(0, 100, size=n) < 5  # outside the scale-invariance range
    maxd[rng_out] *= 0.3
    desc = np.stack([_flip(rng, f["desc"][j], flips) for j in src])
    behind = rng.randint(0, 100, size=n) < 3
    Xw[behind] = (Ow + (Ow - Xw[behind]))  # behind the camera
    return dict(pos=Xw.astype(np.float32), normal=nrm.astype(np.float32), desc=desc, minDist=mind.astype(np.float32),
                maxDist=maxd.astype(np.float32), src=src)


# ---------------------------------------------------------------- SearchByProjection(Current, Last)
@pytest.mark.parametrize("seed,dz,mono,want_mode,check_ori", [(41, 0.9, 0, 1, 1), (42, -0.9, 0, 2, 1), (43, 0.1, 0, 0, 1),
                                                               (44, 0.9, 1, 0, 1), (45, 0.9, 0, 1, 0)])
def test_search_by_projection_last(ref, checker, seed, dz, mono, want_mode, check_ori):
    rng = np.random.RandomState(seed)
    sf, _ = scale_tables()
    nf = 1500
    cur = _features(rng, nf, cluster=(seed == 43))
    Tc = _pose(_rot(0.01, -0.02, 0.005), np.array([0.1, -0.05, 0.3]))
    # last pose: the current camera moved by dz along its optical axis
    Rc, tc = Tc[:3, :3].astype(np.float64), Tc[:3, 3].astype(np.float64)
    Rl = _rot(0.0, 0.01, 0.0) @ Rc
    twc = -Rc.T @ tc
    tl = -Rl @ twc + np.array([0.02, 0.0, dz])
    Tl = _pose(Rl, tl)
    nl = 1400
    P3 = _points_seen_from(rng, cur, Tc, nl, sf, px_noise=3.0)
    last = _features(rng, nl)
    last["angle"] = ((cur["angle"][P3["src"]].astype(np.float64) + rng.normal(0, 6, nl)) % 360.0).astype(np.float32)
    last["octave"] = np.clip(cur["octave"][P3["src"]] + rng.randint(-1, 2, size=nl), 0, 7).astype(np.int32)
    mp_last = np.arange(nl, dtype=np.int32)
    mp_last[rng.randint(0, 100, size=nl) < 10] = -1
    outlier = (rng.randint(0, 100, size=nl) < 5).astype(np.uint8)
    occupied = (rng.randint(0, 100, size=nf) < 5).astype(np.uint8)
    occ = np.nonzero(occupied)[0]
    mp_cur = np.full(nf, -1, np.int32)
    mp_cur[occ] = nl + np.arange(len(occ))
    npts = nl + len(occ)
    nobs = np.concatenate([(rng.randint(0, 100, size=nl) < 85).astype(np.int32), np.ones(len(occ), np.int32)])
    k = Keep()
    pts = k.points(np.concatenate([P3["pos"], np.zeros((len(occ), 3), np.float32)]),
                   np.concatenate([P3["desc"], np.zeros((len(occ), 32), np.uint8)]), nobs=nobs)
    fc = k.feats(cur["x"], cur["y"], cur["angle"], cur["octave"], cur["uright"], cur["desc"], mp=mp_cur, Tcw=Tc)
    fl = k.feats(last["x"], last["y"], last["angle"], last["octave"], last["uright"], last["desc"], mp=mp_last, outlier=outlier,
                 Tcw=Tl)
    m = np.full(nf, -1, np.int32)
    q = np.zeros(nl, proj_query_dtype)
    mode = ctypes.c_int32(-1)
    th = 7.0
    c = cam()
    nr = ref.ref_search_by_projection_last(B(c), B(fc), B(fl), B(pts), th, mono, check_ori, m.ctypes.data, q.ctypes.data, B(mode))
    assert mode.value == want_mode
    keep = np.nonzero(q["octave"] >= 0)[0]
    g = dict(mnMinX=np.float32(0), mnMinY=np.float32(0), mnMaxX=np.float32(W), mnMaxY=np.float32(H), bf=np.float32(BF),
             scale_factors=sf)
    no, mo = checker.search_by_projection_last(np.ascontiguousarray(q[keep]), cur["x"], cur["y"], cur["octave"], cur["angle"],
                                              cur["uright"], occupied, cur["desc"], g, th, mode=mode.value, th_high=100,
                                              check_ori=bool(check_ori))
    assert nr == no and nr > 150
    mo_last = np.where(mo >= 0, keep[np.maximum(mo, 0)], -1)
    assert np.array_equal(m, mo_last)


# ---------------------------------------------------------------- SearchByProjection(Current, KeyFrame*, sAlreadyFound, th, ORBdist)
@pytest.mark.parametrize("seed,th,orb_dist,check_ori,cluster", [(81, 10.0, 100, 1, False), (82, 3.0, 64, 1, False), (83, 10.0, 100, 0, True)])
def test_search_by_projection_relocalisation(ref, checker, seed, th, orb_dist, check_ori, cluster):
    """Tracking::Relocalization's projection search has no kernel of its own: it IS the last-frame search with the window
    levels [l-1, l+1] around the predicted level, every assignment occupying its feature, features that hold any map point
    occupied, no stereo gate and th_high = ORBdist.  The reference's function and that call must agree."""
    rng = np.random.RandomState(seed)
    sf, _ = scale_tables()
    nf, nk = 1500, 1400
    cur = _features(rng, nf, cluster=cluster)
    Tc = _pose(_rot(-0.01, 0.03, 0.002), np.array([-0.2, 0.05, 0.5]))
    P3 = _points_seen_from(rng, cur, Tc, nk, sf, px_noise=3.0)
    kf = _features(rng, nk)
    kf["angle"] = ((cur["angle"][P3["src"]].astype(np.float64) + rng.normal(0, 6, nk)) % 360.0).astype(np.float32)
    mp_kf = np.arange(nk, dtype=np.int32)
    mp_kf[rng.randint(0, 100, size=nk) < 8] = -1
    bad = np.zeros(nk, np.uint8)
    bad[rng.randint(0, 100, size=nk) < 4] = 1
    already = (rng.randint(0, 100, size=nk) < 10).astype(np.uint8)
    occupied = (rng.randint(0, 100, size=nf) < 8).astype(np.uint8)
    occ = np.nonzero(occupied)[0]
    mp_cur = np.full(nf, -1, np.int32)
    mp_cur[occ] = nk + np.arange(len(occ))
    z = len(occ)
    k = Keep()
    pts = k.points(np.concatenate([P3["pos"], np.zeros((z, 3), np.float32)]), np.concatenate([P3["desc"], np.zeros((z, 32), np.uint8)]),
                   bad=np.concatenate([bad, np.zeros(z, np.uint8)]), nobs=np.concatenate([np.ones(nk, np.int32), np.zeros(z, np.int32)]),
                   minDist=np.concatenate([P3["minDist"], np.zeros(z, np.float32)]),
                   maxDist=np.concatenate([P3["maxDist"], np.zeros(z, np.float32)]))
    fc = k.feats(cur["x"], cur["y"], cur["angle"], cur["octave"], cur["uright"], cur["desc"], mp=mp_cur, Tcw=Tc)
    fk = k.feats(kf["x"], kf["y"], kf["angle"], kf["octave"], kf["uright"], kf["desc"], mp=mp_kf, Tcw=Tc)
    m = np.full(nf, -1, np.int32)
    q = np.zeros(nk, proj_query_dtype)
    c = cam()
    nr = ref.ref_search_by_projection_reloc(B(c), B(fc), B(fk), B(pts), already.ctypes.data, th, orb_dist, check_ori,
                                            m.ctypes.data, q.ctypes.data)
    keep = np.nonzero(q["octave"] >= 0)[0]
    assert len(keep) > 0.5 * nk
    g = dict(mnMinX=np.float32(0), mnMinY=np.float32(0), mnMaxX=np.float32(W), mnMaxY=np.float32(H), bf=np.float32(BF),
             scale_factors=sf)
    if getattr(checker, "name", "") == "cuda":  # through the host mirror of the overload
        qq = np.ascontiguousarray(q[keep]).copy()
        qq["invz"], qq["has_obs"] = -5.0, 0  # the wrapper must override both
        no, mo = checker.pkg.ORBmatcher(0.9, bool(check_ori)).SearchByProjectionReloc(qq, cur["x"], cur["y"], cur["octave"],
                                                                                      cur["angle"], occupied, cur["desc"], g, th,
                                                                                      orb_dist)
    else:
        no, mo = checker.search_by_projection_last(np.ascontiguousarray(q[keep]), cur["x"], cur["y"], cur["octave"], cur["angle"],
                                                   np.full(nf, -1, np.float32), occupied, cur["desc"], g, th, mode=0,
                                                   th_high=orb_dist, check_ori=bool(check_ori))
    assert nr == no and nr > 150
    assert np.array_equal(m, np.where(mo >= 0, keep[np.maximum(mo, 0)], -1))


# ---------------------------------------------------------------- SearchForInitialization
@pytest.mark.parametrize("seed,window,nnratio,check_ori,dense", [(91, 100, 0.9, 1, False), (92, 100, 0.9, 0, True), (93, 30, 0.8, 1, False),
                                                                 (94, 100, 0.95, 1, True)])
def test_search_for_initialization(ref, checker, seed, window, nnratio, check_ori, dense):
    """Monocular initialisation matcher: two iterations, the second starting from the first one's vbPrevMatched, as
    Tracking::MonocularInitialization does over consecutive frames.  `dense`: near-duplicate descriptors and many level-0
    features, so that matches get stolen (vMatchedDistance) and the K-lists of the CUDA path run dry."""
    rng = np.random.RandomState(seed)
    sf, _ = scale_tables()
    n1 = n2 = 1500
    f1 = _features(rng, n1)
    if dense:
        f1["octave"] = np.where(rng.randint(0, 100, size=n1) < 75, 0, f1["octave"]).astype(np.int32)
        base = rng.randint(0, 256, size=(6, 32)).astype(np.uint8)
        f1["desc"] = np.stack([_flip(rng, base[rng.randint(0, 6)], 14) for _ in range(n1)])
        f1["x"] = (400 + rng.randint(0, 3000, size=n1) / 10.0).astype(np.float32)
        f1["y"] = (100 + rng.randint(0, 1500, size=n1) / 10.0).astype(np.float32)
    else:
        f1["octave"] = np.where(rng.randint(0, 100, size=n1) < 45, 0, f1["octave"]).astype(np.int32)
    perm = rng.permutation(n1)
    f2 = dict(x=(f1["x"] + rng.normal(8, 6, n1)).astype(np.float32)[perm], y=(f1["y"] + rng.normal(0, 4, n1)).astype(np.float32)[perm],
              octave=np.where(rng.randint(0, 100, size=n1) < 90, f1["octave"], 1).astype(np.int32)[perm],
              angle=((f1["angle"].astype(np.float64) + rng.normal(0, 5, n1)) % 360.0).astype(np.float32)[perm],
              desc=np.stack([_flip(rng, d, 30) for d in f1["desc"]])[perm], uright=np.full(n1, -1, np.float32))
    g = _geom(sf)
    k = Keep()
    F1 = k.feats(f1["x"], f1["y"], f1["angle"], f1["octave"], np.full(n1, -1, np.float32), f1["desc"])
    F2 = k.feats(f2["x"], f2["y"], f2["angle"], f2["octave"], f2["uright"], f2["desc"])
    prev_r = np.ascontiguousarray(np.stack([f1["x"], f1["y"]], 1), np.float32)  # vbPrevMatched starts at F1's keypoints
    prev_c = prev_r.copy()
    c = cam()
    total = 0
    for it in range(2):
        m = np.full(n1, -1, np.int32)
        nr = ref.ref_search_for_initialization(B(c), B(F1), B(F2), prev_r.ctypes.data, window, nnratio, check_ori, m.ctypes.data)
        no, mo = checker.search_for_initialization(prev_c, f1["octave"], f1["angle"], f1["desc"], f2["x"], f2["y"], f2["octave"],
                                                   f2["angle"], f2["desc"], g, window=window, th_low=50, nnratio=nnratio,
                                                   check_ori=bool(check_ori))
        assert nr == no == int((m >= 0).sum())
        assert np.array_equal(m, mo)
        assert np.array_equal(prev_r, prev_c)
        assert (f1["octave"][m >= 0] == 0).all()
        total += nr
    assert total > 150


# ---------------------------------------------------------------- SearchForTriangulation
@pytest.mark.parametrize("seed,only_stereo,check_ori", [(117, 0, 1), (118, 1, 1), (119, 0, 0)])
def test_search_for_triangulation(ref, checker, seed, only_stereo, check_ori):
    d = synth_triangulation(n=1500, seed=seed)
    sf, s2 = scale_tables()
    n = len(d["kf1"]["x"])
    yaw = np.deg2rad(3.0)
    R2 = np.array([[np.cos(yaw), 0, -np.sin(yaw)], [0, 1, 0], [np.sin(yaw), 0, np.cos(yaw)]])
    t2 = -R2 @ np.array([0.6, 0.02, 1.1])
    k = Keep()

    def side(kf, off, T):
        mp = np.where(kf["has_mp"] != 0, off + np.arange(n), -1).astype(np.int32)
        ur = np.where(kf["stereo"] != 0, np.abs(kf["x"]), np.float32(-1.0)).astype(np.float32)
        return k.feats(kf["x"], kf["y"], kf["angle"], kf["octave"], ur, kf["desc"], node=kf["node"], mp=mp, Tcw=T)
    f1 = side(d["kf1"], 0, _pose(np.eye(3), np.zeros(3)))
    f2 = side(d["kf2"], n, _pose(R2, t2))
    pts = k.points(np.zeros((2 * n, 3), np.float32), np.zeros((2 * n, 32), np.uint8))
    F12 = np.ascontiguousarray(d["F12"], np.float32)
    m = np.full(n, -1, np.int32)
    ep = np.zeros(2, np.float32)
    c = cam()
    nr = ref.ref_search_for_triangulation(B(c), B(f1), B(f2), B(pts), F12.ctypes.data, only_stereo, check_ori, m.ctypes.data,
                                          ep.ctypes.data)
    assert abs(float(ep[0]) - float(d["ex"])) < 0.05 and abs(float(ep[1]) - float(d["ey"])) < 0.05
    no, mo = checker.search_for_triangulation(d["kf1"], d["kf2"], F12, float(ep[0]), float(ep[1]), sf, s2,
                                             only_stereo=bool(only_stereo), check_ori=bool(check_ori))
    assert nr == no and nr > 100
    assert np.array_equal(m, mo)


# ---------------------------------------------------------------- Fuse x2, SearchByProjection(KeyFrame*, Scw), SearchBySim3
def _keyframe_world(seed, nf=1500, nq=1800, cluster=False, S=None):
    rng = np.random.RandomState(seed)
    sf, s2 = scale_tables()
    f = _features(rng, nf, cluster=cluster)
    T = _pose(_rot(0.02, 0.03, -0.01), np.array([0.3, 0.1, -0.2]))
    P3 = _points_seen_from(rng, f, T, nq, sf, S=S)
    return rng, sf, s2, f, T, P3


def _geom(sf):
    return dict(mnMinX=np.float32(0), mnMinY=np.float32(0), mnMaxX=np.float32(W), mnMaxY=np.float32(H), bf=np.float32(BF),
                scale_factors=sf)


def _sim3(rng):
    s = 1.07
    R = _rot(0.02, 0.03, -0.01)
    S = np.eye(4)
    S[:3, :3] = s * R
    S[:3, 3] = s * np.array([0.3, 0.1, -0.2])
    return S.astype(np.float32)


@pytest.mark.parametrize("seed,cluster,use_scw,th", [(51, False, False, 3.0), (52, True, False, 3.0), (53, False, True, 4.0),
                                                     (54, True, True, 6.0)])
def test_fuse(ref, checker, seed, cluster, use_scw, th):
    S = _sim3(None) if use_scw else None
    rng, sf, s2, f, T, P3 = _keyframe_world(seed, cluster=cluster, S=S)
    nf, nq = len(f["x"]), len(P3["pos"])
    # keyframe features that already hold a point (own points: nq .. nq+nown-1); some queries ARE own points / bad / NULL
    own = np.nonzero(rng.randint(0, 100, size=nf) < 40)[0]
    mp = np.full(nf, -1, np.int32)
    mp[own] = nq + np.arange(len(own))
    npts = nq + len(own)
    bad = np.zeros(npts, np.uint8)
    bad[:nq] = rng.randint(0, 100, size=nq) < 4
    nobs = rng.randint(1, 6, size=npts).astype(np.int32)
    qp = np.arange(nq, dtype=np.int32)
    if not use_scw:  # Fuse(pKF, vpMapPoints) tolerates NULL entries (:1042); the Scw overload takes a list without them
        qp[rng.randint(0, 100, size=nq) < 3] = -1
    dup = np.nonzero(rng.randint(0, 100, size=nq) < 3)[0]
    qp[dup] = nq + rng.randint(0, len(own), size=len(dup))  # points the keyframe already observes
    z3 = np.zeros((len(own), 3), np.float32)

    def ext(a, fill):
        return np.concatenate([a, np.full((len(own),) + a.shape[1:], fill, a.dtype)])
    k = Keep()
    pts = k.points(ext(P3["pos"], 0), ext(P3["desc"], 0), normal=ext(P3["normal"], 0), bad=bad, nobs=nobs,
                   minDist=ext(P3["minDist"], 0), maxDist=ext(P3["maxDist"], 0))
    kf = k.feats(f["x"], f["y"], f["angle"], f["octave"], f["uright"], f["desc"], mp=mp, Tcw=T)
    best = np.full(nq, -1, np.int32)
    q = np.zeros(nq, win_query_dtype)
    c = cam()
    Sp = None if S is None else np.ascontiguousarray(S.reshape(16))
    nr = ref.ref_fuse(B(c), B(kf), B(pts), qp.ctypes.data, nq, None if Sp is None else Sp.ctypes.data, th, best.ctypes.data,
                      q.ctypes.data)
    assert q["valid"].sum() > nq * 0.5
    no, bo, _ = checker.search_windows(q, f["x"], f["y"], f["octave"], f["uright"], (np.float32(1.0) / s2).astype(np.float32), None,
                                      f["desc"], _geom(sf), chi2=not use_scw, greedy=False, th_dist=50)
    assert nr == no and nr > 150
    assert np.array_equal(best, bo)


@pytest.mark.parametrize("seed,cluster,th", [(61, False, 10), (62, True, 10), (63, False, 4)])
def test_search_by_projection_scw(ref, checker, seed, cluster, th):
    S = _sim3(None)
    rng, sf, s2, f, T, P3 = _keyframe_world(seed, cluster=cluster, S=S)
    nf, nq = len(f["x"]), len(P3["pos"])
    occupied = (rng.randint(0, 100, size=nf) < 15).astype(np.uint8)
    occ = np.nonzero(occupied)[0]
    matched_in = np.full(nf, -1, np.int32)
    # already-matched features hold either an outside point or one of the query points (which is then skipped)
    matched_in[occ] = nq + np.arange(len(occ))
    some = occ[::5]
    matched_in[some] = rng.choice(nq, size=len(some), replace=False)
    npts = nq + len(occ)
    bad = np.zeros(npts, np.uint8)
    bad[:nq] = rng.randint(0, 100, size=nq) < 4

    def ext(a, fill):
        return np.concatenate([a, np.full((len(occ),) + a.shape[1:], fill, a.dtype)])
    k = Keep()
    pts = k.points(ext(P3["pos"], 0), ext(P3["desc"], 0), normal=ext(P3["normal"], 0), bad=bad, minDist=ext(P3["minDist"], 0),
                   maxDist=ext(P3["maxDist"], 0))
    kf = k.feats(f["x"], f["y"], f["angle"], f["octave"], f["uright"], f["desc"], Tcw=T)
    qp = np.arange(nq, dtype=np.int32)
    best = np.full(nq, -1, np.int32)
    q = np.zeros(nq, win_query_dtype)
    Sp = np.ascontiguousarray(S.reshape(16))
    c = cam()
    nr = ref.ref_search_by_projection_scw(B(c), B(kf), B(pts), qp.ctypes.data, nq, matched_in.ctypes.data, Sp.ctypes.data, th,
                                          best.ctypes.data, q.ctypes.data)
    no, bo, _ = checker.search_windows(q, f["x"], f["y"], f["octave"], f["uright"], None, occupied, f["desc"], _geom(sf),
                                      chi2=False, greedy=True, th_dist=50)
    assert nr == no and nr > 150
    assert np.array_equal(best, bo)


@pytest.mark.parametrize("seed,th", [(71, 7.5), (72, 3.0)])
def test_search_by_sim3(ref, checker, seed, th):
    rng = np.random.RandomState(seed)
    sf, s2 = scale_tables()
    n = 1200
    # keyframe 2 sees the world at scale s; S12 maps camera-2 coordinates into camera-1 coordinates
    T1 = _pose(_rot(0.01, 0.02, 0.0), np.array([0.1, 0.0, 0.2]))
    T2 = _pose(_rot(-0.02, 0.05, 0.01), np.array([-0.4, 0.05, 0.3]))
    f1 = _features(rng, n)
    P1 = _points_seen_from(rng, f1, T1, n, sf, px_noise=0.3, flips=1, src=np.arange(n))  # kf1's own points, on its features
    s12 = 1.05
    R1, t1 = T1[:3, :3].astype(np.float64), T1[:3, 3].astype(np.float64)
    R2, t2 = T2[:3, :3].astype(np.float64), T2[:3, 3].astype(np.float64)
    R12 = R1 @ R2.T @ _rot(0.002, -0.003, 0.001)
    t12 = (t1 - R12 @ t2) * 1.0 + np.array([0.01, -0.01, 0.02])
    # kf2's features: where kf1's points land in camera 2 through S21, plus noise; kf2's points: its own features' points
    X1 = P1["pos"].astype(np.float64)
    Xc1 = (R1 @ X1.T).T + t1
    sR21 = (1.0 / s12) * R12.T
    t21 = -sR21 @ t12
    Xc2 = (sR21 @ Xc1.T).T + t21
    perm = rng.permutation(n)
    u2 = FX * Xc2[:, 0] / Xc2[:, 2] + CX + rng.normal(0, 1.5, n)
    v2 = FY * Xc2[:, 1] / Xc2[:, 2] + CY + rng.normal(0, 1.0, n)
    f2 = dict(x=u2.astype(np.float32)[perm], y=v2.astype(np.float32)[perm],
              octave=np.clip(f1["octave"] + rng.randint(-1, 1, size=n), 0, 7).astype(np.int32)[perm],
              angle=f1["angle"][perm], uright=np.full(n, -1, np.float32),
              desc=np.stack([_flip(rng, d, 25) for d in f1["desc"]])[perm])
    # kf2's map points: the same physical points expressed in kf2's world: Xw2 = R2^T (Xc2 - t2)
    X2 = (R2.T @ (Xc2 - t2).T).T[perm]
    d2 = np.linalg.norm(Xc2, axis=1)[perm]
    max2 = d2 * sf[np.clip(f2["octave"] + rng.randint(0, 2, size=n), 0, 7)]
    d1 = np.linalg.norm(Xc1, axis=1)
    P1["maxDist"] = (d1 * sf[np.clip(f1["octave"] + rng.randint(0, 2, size=n), 0, 7)]).astype(np.float32)
    P1["minDist"] = (P1["maxDist"] / sf[7]).astype(np.float32)
    pos = np.concatenate([P1["pos"], X2.astype(np.float32)])
    desc = np.concatenate([np.stack([_flip(rng, d, 10) for d in f1["desc"]]), np.stack([_flip(rng, d, 10) for d in f2["desc"]])])
    maxD = np.concatenate([P1["maxDist"], max2.astype(np.float32)])
    minD = (maxD / sf[7]).astype(np.float32)
    bad = (rng.randint(0, 100, size=2 * n) < 4).astype(np.uint8)
    mp1 = np.where(rng.randint(0, 100, size=n) < 85, np.arange(n), -1).astype(np.int32)
    mp2 = np.where(rng.randint(0, 100, size=n) < 85, n + np.arange(n), -1).astype(np.int32)
    # matches found earlier (by BoW): kf1 feature i <-> kf2 feature inv[i]
    inv = np.argsort(perm).astype(np.int32)
    prior = np.full(n, -1, np.int32)
    pick = np.nonzero((rng.randint(0, 100, size=n) < 15) & (mp2[inv] >= 0))[0]
    prior[pick] = inv[pick]
    k = Keep()
    pts = k.points(pos, desc, bad=bad, minDist=minD, maxDist=maxD)
    kf1 = k.feats(f1["x"], f1["y"], f1["angle"], f1["octave"], f1["uright"], f1["desc"], mp=mp1, Tcw=T1)
    kf2 = k.feats(f2["x"], f2["y"], f2["angle"], f2["octave"], f2["uright"], f2["desc"], mp=mp2, Tcw=T2)
    R12f = np.ascontiguousarray(R12.reshape(9), np.float32)
    t12f = np.ascontiguousarray(t12, np.float32)
    m12 = np.full(n, -1, np.int32)
    q12 = np.zeros(n, win_query_dtype)
    q21 = np.zeros(n, win_query_dtype)
    c = cam()
    nr = ref.ref_search_by_sim3(B(c), B(kf1), B(kf2), B(pts), prior.ctypes.data, s12, R12f.ctypes.data, t12f.ctypes.data, th,
                                m12.ctypes.data, q12.ctypes.data, q21.ctypes.data)
    assert q12["valid"].sum() > n * 0.4 and q21["valid"].sum() > n * 0.4
    _, b12, _ = checker.search_windows(q12, f2["x"], f2["y"], f2["octave"], f2["uright"], None, None, f2["desc"], _geom(sf),
                                      th_dist=100)
    _, b21, _ = checker.search_windows(q21, f1["x"], f1["y"], f1["octave"], f1["uright"], None, None, f1["desc"], _geom(sf),
                                      th_dist=100)
    want = np.full(n, -1, np.int32)
    for i1 in range(n):
        i2 = b12[i1]
        if i2 >= 0 and b21[i2] == i1:
            want[i1] = i2
    assert nr == int((want >= 0).sum()) and nr > 100
    assert np.array_equal(m12, want)
