"""Pins the oracle's distinctive-descriptor selection against the REFERENCE'S OWN MapPoint: /root/reference/src/MapPoint.cc
compiled in place, unmodified, with its own MapPoint.h (oracle/_ref/libref_mappoint.so; KeyFrame / Frame / Map are
stand-ins).  Observations are added in random order; the reference iterates them in std::map<KeyFrame*, size_t> order
(keyframe address = keyframe id here), skips bad keyframes, and keeps the descriptor with the least median Hamming
distance to the others (src/MapPoint.cc:359-440) — the oracle must choose the same one."""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_mappoint.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")
vp, c_i, c_f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


@pytest.mark.parametrize("seed,near_dup", [(1, False), (2, True), (3, True)])
def test_distinctive_descriptor_equals_reference_mappoint(checker, seed, near_dup):
    rng = np.random.RandomState(seed)
    n_kf, per_kf, n_pts = 40, 300, 600
    kf_off = (np.arange(n_kf + 1) * per_kf).astype(np.int32)
    base = rng.randint(0, 256, size=(n_pts, 32)).astype(np.uint8)
    kf_desc = rng.randint(0, 256, size=(n_kf * per_kf, 32)).astype(np.uint8)
    kf_oct = rng.randint(0, 8, size=n_kf * per_kf).astype(np.int32)
    kf_ow = rng.uniform(-5, 5, size=(n_kf, 3)).astype(np.float32)
    kf_bad = (rng.randint(0, 100, size=n_kf) < 12).astype(np.uint8)
    offsets, obs_kf, obs_idx = [0], [], []
    used = np.zeros(n_kf, np.int32)
    for p in range(n_pts):
        k = int(rng.choice([0, 1, 2, 3, 5, 8, 12, 20]))
        kfs = rng.choice(n_kf, size=k, replace=False)
        for kf in kfs:
            i = used[kf]
            used[kf] += 1
            d = base[p].copy()
            flips = int(rng.randint(0, 3 if near_dup else 60))  # near-duplicates: many equal medians -> first minimum wins
            for b in rng.choice(256, size=flips, replace=False):
                d[b >> 3] ^= np.uint8(1 << (b & 7))
            kf_desc[kf * per_kf + i] = d
            obs_kf.append(kf)
            obs_idx.append(i)
        offsets.append(len(obs_kf))
    offsets = np.array(offsets, np.int32)
    obs_kf, obs_idx = np.array(obs_kf, np.int32), np.array(obs_idx, np.int32)
    pos = rng.uniform(-20, 20, size=(n_pts, 3)).astype(np.float32)
    out_desc = np.zeros((n_pts, 32), np.uint8)
    out_n, out_mm = np.zeros((n_pts, 3), np.float32), np.zeros((n_pts, 2), np.float32)
    R = ctypes.CDLL(LIB)
    R.ref_mappoint_update.argtypes = [c_i, vp, vp, vp, vp, vp, c_i, vp, vp, vp, vp, c_i, c_f, vp, vp, vp]
    R.ref_mappoint_update.restype = None
    R.ref_mappoint_update(n_kf, kf_off.ctypes.data, kf_desc.ctypes.data, kf_oct.ctypes.data, kf_ow.ctypes.data, kf_bad.ctypes.data,
                          n_pts, offsets.ctypes.data, obs_kf.ctypes.data, obs_idx.ctypes.data, pos.ctypes.data, 8, 1.2,
                          out_desc.ctypes.data, out_n.ctypes.data, out_mm.ctypes.data)
    # what the caller of the oracle / the CUDA path passes: per point, descriptors in keyframe-id order, bad keyframes removed
    lists, offs = [], [0]
    for p in range(n_pts):
        q = np.arange(offsets[p], offsets[p + 1])
        q = q[np.argsort(obs_kf[q], kind="stable")]
        q = q[kf_bad[obs_kf[q]] == 0]
        lists.append(kf_desc[obs_kf[q] * per_kf + obs_idx[q]])
        offs.append(offs[-1] + len(q))
    flat = np.concatenate([l for l in lists if len(l)]) if offs[-1] else np.zeros((0, 32), np.uint8)
    best = checker.distinctive_descriptors(flat, np.array(offs, np.int32))
    n_checked = 0
    for p in range(n_pts):
        if len(lists[p]) == 0:
            assert best[p] == -1 and not out_desc[p].any()  # the reference returns before assigning (:381-382)
        else:
            assert np.array_equal(out_desc[p], lists[p][best[p]]), p
            n_checked += 1
    assert n_checked > 0.7 * n_pts
    # UpdateNormalAndDepth (:476-520): mean of the unit viewing rays of the non-... all observing keyframes
    p = int(np.argmax(np.diff(offsets)))
    q = np.arange(offsets[p], offsets[p + 1])
    rays = pos[p].astype(np.float64) - kf_ow[obs_kf[q]].astype(np.float64)
    want = (rays / np.linalg.norm(rays, axis=1)[:, None]).mean(axis=0)
    assert np.allclose(out_n[p], want, atol=1e-5)
