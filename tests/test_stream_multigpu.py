"""Rank-sharded stream == single-GPU stream (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_stream_multigpu.py -m gpu`).

A global sequence of 2 x F frames is processed (a) by two ranks over NCCL, each extracting / stereo-matching its own F
frames and matching its first frame against the other rank's last one through the all-gathered record, for both exchange
modes (shard-boundary record only, and every left-image record as in SURVEY.md §8e), and (b) by ONE StereoStream over
all 2F frames.  Keypoint records, descriptors, stereo matches and temporal matches must be identical."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from synth import synth_stereo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["boundary", "all"])
def test_two_ranks_equal_one_gpu(pkg, tmp_path, exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    F, world, w, h = 3, 2, 1241, 376
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tools", "multi_gpu_check.py"),
                           str(F), str(tmp_path), exchange], env=env, timeout=600)
    stream_mod = importlib.import_module("self_commit_orb-slam2_b200.stream")
    G = F * world
    pairs = [synth_stereo(w, h, 700 + i) for i in range(G)]
    imgs = np.stack([p[0] for p in pairs] + [p[1] for p in pairs])
    ss = stream_mod.StereoStream(G, w, h, 2000, stereo=True, project=True)
    ss.upload(torch.from_numpy(imgs))
    ss.step_device()
    torch.cuda.synchronize()
    ss.ex.check()
    counts = ss.counts[1:1 + G].cpu().numpy()
    kps, desc = ss.kps[1:1 + G].cpu().numpy(), ss.desc[1:1 + G].cpu().numpy()
    nmatch, match = ss.nmatch.cpu().numpy(), ss.match.cpu().numpy()
    ur, ns = ss.ur.cpu().numpy(), ss.nstereo.cpu().numpy()
    pm, npm = ss.pmatch.cpu().numpy(), ss.npmatch.cpu().numpy()
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        lo = r * F
        assert np.array_equal(d["counts"][:F], counts[lo:lo + F])
        for f in range(F):
            n = counts[lo + f]
            assert np.array_equal(d["kps"][f, :n], kps[lo + f, :n]) and np.array_equal(d["desc"][f, :n], desc[lo + f, :n])
            assert np.array_equal(d["uright"][f, :n], ur[lo + f, :n])
            assert np.array_equal(d["match"][f, :n], match[lo + f, :n])  # frame lo+f against its global predecessor
            assert np.array_equal(d["pmatch"][f, :n], pm[lo + f, :n])   # SearchByProjection(Cur, Last): the predecessor's
                                                                         # depths crossed the rank boundary with its record
        assert np.array_equal(d["nmatch"], nmatch[lo:lo + F]) and np.array_equal(d["nstereo"], ns[lo:lo + F])
        assert np.array_equal(d["npmatch"], npm[lo:lo + F]) and npm.sum() > 0
