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
