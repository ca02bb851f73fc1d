"""Breakdown of one end-to-end (host-buffer) stream step: whole step without LocalBA, extraction call alone."""
import sys, time, importlib, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import bench
pkg = importlib.import_module('self_commit_orb-slam2_b200')
sm = importlib.import_module('self_commit_orb-slam2_b200.stream')
F = 160
ss = sm.StereoStream(F, bench.W_IMG, bench.H_IMG, bench.NFEAT, ba_problem=bench.ba_window(), ba_every=5)
imgs = torch.from_numpy(bench.make_images(16, F)).pin_memory().numpy()
ss.step_host(imgs, run_ba=False)
L = pkg.lib(); _vp = ctypes.c_void_p
for rep in range(3):
    t0 = time.perf_counter()
    ss.step_host(imgs, run_ba=False)
    t1 = time.perf_counter()
    pkg._check(L.b2s_extract_batch(ss.ex._h, ctypes.cast(ss._h_ptrs, _vp), 2 * F, ss.w, ss.h, ss.w, ss._h_kps[1:].ctypes.data_as(_vp),
                                   ss._h_desc[1:].ctypes.data_as(_vp), ss.cap, ss._h_n[1:].ctypes.data_as(_vp)))
    t2 = time.perf_counter()
    ss.step_host(imgs, run_ba=True)
    t3 = time.perf_counter()
    print("step_host(no BA) %.2f ms   extract_batch alone %.2f ms   step_host(with BA, serial) %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
