"""Per-kernel view of the extractor alone (run through gpurun, optionally under ncu): B images of 1241x376, resident."""
import importlib
import sys
import os
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
pkg = importlib.import_module("self_commit_orb-slam2_b200")
from synth import synth_stereo
B = int(sys.argv[1]) if len(sys.argv) > 1 else 320
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H = 1241, 376
base = [synth_stereo(W, H, i)[0] for i in range(16)]
imgs = np.stack([base[i % 16] for i in range(B)])
d = torch.from_numpy(imgs).cuda()
ex = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
cap = ex.cap
kps = torch.zeros((B, cap, 28), dtype=torch.uint8, device="cuda")
desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(B, dtype=torch.int32, device="cuda")
st = torch.cuda.Stream()
for r in range(reps):
    torch.cuda.synchronize()
    t = time.perf_counter()
    ex.extract_batch_device(d.data_ptr(), W * H, B, W, H, W, kps.data_ptr(), desc.data_ptr(), cnt.data_ptr(), cap, stream=st.cuda_stream)
    torch.cuda.synchronize()
    print("rep %d: %.3f ms, mean keypoints %.1f" % (r, (time.perf_counter() - t) * 1e3, cnt.float().mean().item()))
