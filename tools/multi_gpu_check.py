"""Rank script of tests/test_stream_multigpu.py: every rank runs one stream step on ITS shard of a global frame sequence
(frame g = global index, images seeded by g) and writes what it produced; the test compares the concatenation with a
single-GPU run over the whole sequence.  usage: torchrun --nproc-per-node N tools/multi_gpu_check.py F_PER_RANK OUT_DIR EXCHANGE"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
from synth import synth_stereo

F = int(sys.argv[1])
out_dir = sys.argv[2]
exchange = sys.argv[3] if len(sys.argv) > 3 else "boundary"
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream_mod = importlib.import_module("self_commit_orb-slam2_b200.stream")
w, h = 1241, 376
pairs = [synth_stereo(w, h, 700 + rank * F + i) for i in range(F)]
imgs = np.stack([p[0] for p in pairs] + [p[1] for p in pairs])
ss = stream_mod.StereoStream(F, w, h, 2000, device=local, rank=rank, world=world, exchange=exchange, stereo=True, project=True)
ss.upload(torch.from_numpy(imgs))
ss.step_device()
torch.cuda.synchronize()
ss.ex.check()
np.savez(os.path.join(out_dir, "rank%d.npz" % rank), counts=ss.counts[1:].cpu().numpy(), kps=ss.kps[1:1 + F].cpu().numpy(),
         desc=ss.desc[1:1 + F].cpu().numpy(), nmatch=ss.nmatch.cpu().numpy(), match=ss.match.cpu().numpy(),
         uright=ss.ur.cpu().numpy(), nstereo=ss.nstereo.cpu().numpy(), pmatch=ss.pmatch.cpu().numpy(),
         npmatch=ss.npmatch.cpu().numpy())
dist.barrier()
dist.destroy_process_group()
