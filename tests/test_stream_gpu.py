"""Batched stream mode on one GPU: device-resident step == host-buffer step == oracle, frame by frame."""
import ctypes
import importlib

import numpy as np
import pytest

from synth import synth_stereo, synth_local_ba

pytestmark = pytest.mark.gpu


def test_stream_step_matches_oracle(pkg, oracle):
    import torch
    stream_mod = importlib.import_module("self_commit_orb-slam2_b200.stream")
    F, w, h = 4, 1241, 376
    pairs = [synth_stereo(w, h, 50 + i) for i in range(F)]
    imgs = np.stack([p[0] for p in pairs] + [p[1] for p in pairs])
    ba = synth_local_ba(n_kf=8, n_fixed=2, n_mp=200, obs_per_mp=4, seed=5)
    ss = stream_mod.StereoStream(F, w, h, 2000, ba_problem=ba, ba_every=2)
    ss.upload(torch.from_numpy(imgs))
    ba_out = ss.step_device()
    torch.cuda.synchronize()
    ss.ex.check()
    counts_dev = ss.counts.cpu().numpy()
    nm_dev = ss.nmatch.cpu().numpy()
    match_dev = ss.match.cpu().numpy()
    n_host, nm_host, ba_host, (kps_h, desc_h, match_h) = ss.step_host(imgs)
    assert np.array_equal(counts_dev[1:], n_host)
    assert np.array_equal(nm_dev, nm_host)
    assert np.array_equal(match_dev, match_h)
    # device records == host records
    kd = ss.kps[1:].cpu().numpy().view(pkg.keypoint_dtype).reshape(2 * F, ss.cap)
    dd = ss.desc[1:].cpu().numpy()
    for b in range(2 * F):
        assert np.array_equal(kd[b, :n_host[b]], kps_h[b, :n_host[b]])
        assert np.array_equal(dd[b, :n_host[b]], desc_h[b, :n_host[b]])
    # oracle: same ring matching on the CPU
    oe = oracle.extractor(2000, 1.2, 8, 20, 7)
    ref = [oe(imgs[i]) for i in range(2 * F)]
    for b in range(2 * F):
        assert len(ref[b][0]) == n_host[b]
        assert np.array_equal(ref[b][1], desc_h[b, :n_host[b]])
    for f in range(F):
        a = (f - 1) % F
        ka, da = ref[a]
        kb, db = ref[f]
        on, om = oracle.search_by_bow(da, np.zeros(len(da), np.int32), np.ones(len(da), np.uint8), ka["angle"].copy(),
                                      db, np.zeros(len(db), np.int32), kb["angle"].copy(), nnratio=0.7)
        assert on == nm_host[f]
        assert np.array_equal(om, match_h[f, :len(db)])
    ref_ba = oracle.local_ba(ba)
    for out in (ba_out[0], ba_host[-1]):
        assert out["n_trials"] == ref_ba["n_trials"]
        assert np.array_equal(out["outlier"], ref_ba["outlier"])
    assert ss.launch_count() > 0


@pytest.mark.parametrize("motion", [None, "backward", "slow"])
def test_stream_projection_matches_per_frame_api_and_oracle(pkg, oracle, motion):
    """The batched SearchByProjection(CurrentFrame, LastFrame) of the stream: queries built on the device == their numpy
    restatement (bit for bit), matches == the one-pair C entry point == the oracle (pinned to src/ORBmatcher.cc by
    test_oracle_reference_matcher), host-buffer sequence entry == device path."""
    import torch
    stream_mod = importlib.import_module("self_commit_orb-slam2_b200.stream")
    F, w, h = 3, 1241, 376
    pairs = [synth_stereo(w, h, 90 + i) for i in range(F)]
    imgs = np.stack([p[0] for p in pairs] + [p[1] for p in pairs])
    T = None
    if motion == "backward":
        T = np.array([[1, 0, 0, 0.0], [0, 1, 0, 0.01], [0, 0, 1, 0.9]], np.float32)
    elif motion == "slow":
        T = np.array([[1, 0, 0, 0.05], [0, 1, 0, 0.0], [0, 0, 1, -0.1]], np.float32)
    ss = stream_mod.StereoStream(F, w, h, 2000, stereo=True, project=True, motion=T)
    assert ss.proj_mode == {None: 1, "backward": 2, "slow": 0}[motion]
    ss.upload(torch.from_numpy(imgs))
    ss.step_device()
    torch.cuda.synchronize()
    ss.ex.check()
    cap = ss.cap
    counts = ss.counts.cpu().numpy()
    kps = ss.kps.cpu().numpy().view(pkg.keypoint_dtype).reshape(1 + 2 * F, cap)
    desc = ss.desc.cpu().numpy()
    dp, ur = ss.dp_all.cpu().numpy(), ss.ur.cpu().numpy()
    q_dev = ss.q.cpu().numpy().view(pkg.proj_query_dtype).reshape(F, cap)
    nq, pm, npm = ss.nq.cpu().numpy(), ss.pmatch.cpu().numpy(), ss.npmatch.cpu().numpy()
    assert np.array_equal(nq, counts[:F])
    assert (dp[0] == dp[F]).all()
    q_host = ss.track_queries_host(kps[:F], desc[:F], dp[:F], counts[:F], np.tile(ss.motion.reshape(1, 12), (F, 1)), ss.fx, ss.fy,
                                   ss.cx, ss.cy, 1)
    m1 = pkg.ORBmatcher(0.9, True, max_features=cap)
    total = 0
    for f in range(F):
        n_last, n_cur = counts[f], counts[1 + f]
        assert q_dev[f, :n_last].tobytes() == q_host[f, :n_last].tobytes()
        assert (q_dev[f, :n_last]["invz"] >= 0).sum() > 200  # the stereo points of the last frame
        k = kps[1 + f, :n_cur]
        args = (np.ascontiguousarray(q_dev[f, :n_last]), k["x"].copy(), k["y"].copy(), k["octave"].copy(), k["angle"].copy(),
                ur[f, :n_cur].copy(), None, desc[1 + f, :n_cur].copy(), ss.geom, ss.proj_th)
        nm_one, match_one = m1.SearchByProjection(*args, mode=ss.proj_mode)
        assert nm_one == npm[f]
        assert np.array_equal(match_one, pm[f, :n_cur])
        assert (pm[f, n_cur:] == -1).all()
        nm_o, match_o = oracle.search_by_projection_last(*args, mode=ss.proj_mode)
        assert nm_o == npm[f]
        assert np.array_equal(match_o, pm[f, :n_cur])
        total += nm_one
    assert total > 0
    # host-buffer step (b2s_search_by_projection_sequence) == device-resident step
    ss.step_host(imgs)
    assert np.array_equal(ss._h_npm, npm)
    assert np.array_equal(ss._h_pm, pm)
    # host-buffer batch entry with caller-made queries
    mb = pkg.ORBmatcher(0.9, True, max_features=cap, max_batch=F)
    nm_b, match_b = mb.SearchByProjectionBatch(q_dev, nq, kps[1:1 + F], ur, desc[1:1 + F], counts[1:1 + F], ss.geom, ss.proj_th,
                                               mode=ss.proj_mode)
    assert np.array_equal(nm_b, npm) and np.array_equal(match_b, pm)
