"""Host-side logic of the stream's motion-model matcher (no GPU): the numpy restatement of b2s_track_queries_device against an
independent float64 computation, and the forward / backward rule of src/ORBmatcher.cc:1589-1598."""
import importlib

import numpy as np


def _mod(pkg):
    return importlib.import_module("self_commit_orb-slam2_b200.stream")


def test_projection_mode_rule(pkg):
    S = _mod(pkg).StereoStream
    mb = 386.1448 / 718.856
    I = np.eye(3)
    fwd = np.hstack([I, [[0.0], [0.0], [-0.8]]])   # points come 0.8 m closer: the camera moved forward
    back = np.hstack([I, [[0.0], [0.0], [0.9]]])
    slow = np.hstack([I, [[0.1], [0.0], [-0.2]]])
    assert S.projection_mode(fwd, mb) == 1 and S.projection_mode(back, mb) == 2 and S.projection_mode(slow, mb) == 0
    # tlc = -R^T t: a rotated camera
    a = 0.3
    R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    t = np.array([0.0, 0.0, -1.0])
    assert S.projection_mode(np.hstack([R, t[:, None]]), mb) == (1 if (-R.T @ t)[2] > mb else 0)


def test_track_queries_host_matches_float64(pkg):
    S = _mod(pkg).StereoStream
    rng = np.random.RandomState(4)
    B, cap = 3, 64
    kps = np.zeros((B, cap), pkg.keypoint_dtype)
    kps["x"] = rng.uniform(0, 1241, size=(B, cap)).astype(np.float32)
    kps["y"] = rng.uniform(0, 376, size=(B, cap)).astype(np.float32)
    kps["angle"] = rng.uniform(0, 360, size=(B, cap)).astype(np.float32)
    kps["octave"] = rng.randint(0, 8, size=(B, cap))
    desc = rng.randint(0, 256, size=(B, cap, 32)).astype(np.uint8)
    depth = np.where(rng.rand(B, cap) < 0.5, rng.uniform(2, 60, size=(B, cap)), -1).astype(np.float32)
    depth[0, 0] = 0.3  # ends up behind the camera after the motion below
    n = np.array([cap, 40, 0], np.int32)
    a = 0.02
    T = np.array([[np.cos(a), 0, np.sin(a), 0.05], [0, 1, 0, -0.01], [-np.sin(a), 0, np.cos(a), -0.8]], np.float32)
    fx, fy, cx, cy = 718.856, 718.856, 607.1928, 185.2157
    q = S.track_queries_host(kps, desc, depth, n, np.tile(T.reshape(1, 12), (B, 1)), fx, fy, cx, cy, 1)
    assert q.dtype == pkg.proj_query_dtype and q.shape == (B, cap)
    for b in range(B):
        for i in range(cap):
            if i >= n[b]:
                assert q[b, i].tobytes() == bytes(pkg.proj_query_dtype.itemsize)
                continue
            assert q[b, i]["octave"] == kps[b, i]["octave"] and q[b, i]["angle"] == kps[b, i]["angle"]
            assert np.array_equal(q[b, i]["desc"], desc[b, i]) and q[b, i]["has_obs"] == 1
            z = float(depth[b, i])
            if not z > 0:
                assert q[b, i]["invz"] == -1 and q[b, i]["u"] == 0 and q[b, i]["v"] == 0
                continue
            X = np.array([(float(kps[b, i]["x"]) - cx) * z / fx, (float(kps[b, i]["y"]) - cy) * z / fy, z])
            Xc = T[:, :3].astype(np.float64) @ X + T[:, 3]
            if 1.0 / Xc[2] < 0:
                assert q[b, i]["invz"] == -1
                continue
            assert abs(q[b, i]["invz"] - 1.0 / Xc[2]) <= 1e-5 / abs(Xc[2])
            assert abs(q[b, i]["u"] - (fx * Xc[0] / Xc[2] + cx)) < 2e-2 and abs(q[b, i]["v"] - (fy * Xc[1] / Xc[2] + cy)) < 2e-2
    assert (q[0]["invz"] == -1).sum() > 0 and (q[0]["invz"] > 0).sum() > 0
