#!/usr/bin/env python3
"""bench.py — stereo frames/s of the hot path (extract + match + LocalBA) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                 # this framework (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus 1 --steps K --warmup W  # the reference algorithm on the host cores (oracle)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one batch of F synthetic KITTI-shaped stereo frames (1241x376, 2000 features, 8 levels) per GPU:
extraction of the 2F images, Frame::ComputeStereoMatches of the F pairs, temporal SearchByBoW (2000x2000 brute force, one
vocabulary node) of every left image against its predecessor, Optimizer::PoseOptimization of every frame and one LocalBA
per 5 frames (32 DIFFERENT windows per step, 30-60 keyframes); with N > 1 the left-image feature records are all-gathered
over NCCL.  Frames are distinct (seed = frame index).  Prints ONE JSON line (see DESIGN.md "Measurement").
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W_IMG, H_IMG, NFEAT = 1241, 376, 2000
FX, FY, CX, CY, BF = 718.856, 718.856, 607.1928, 185.2157, 386.1448  # KITTI 00-02 (Examples/Stereo/KITTI00-02.yaml)
PROJ_TH = 7.0  # SearchByProjection(CurrentFrame, LastFrame) window for stereo (src/Tracking.cc:892-897)
# pose of the current camera relative to the last one (3 x 4): 0.8 m forward with a small yaw -> the reference's forward branch
STREAM_MOTION = np.array([[np.cos(0.01), 0, np.sin(0.01), 0.02], [0, 1, 0, 0.0], [-np.sin(0.01), 0, np.cos(0.01), -0.8]], np.float32)
# algorithmic bytes per 1241x376 image
B_STAGE_IMAGE = 9359539          # all extractor stages
B_TILE_IMAGE = 2385248 + 1444097 + 2888194  # the stages the fused tile kernel replaces: pyramid R+W, FAST read, blur R+W
STAGE_NAMES = ["resize_chain(legacy)", "tile_fast_blur_pyramid+cells", "quadtree", "blur(legacy)", "orient_describe"]
# dram__bytes_read.sum + dram__bytes_write.sum of the eight k_tile launches of a 320-image batch (ncu --set full), per image;
# None until a capture of the current kernel is committed under profiles/
TRAFFIC_TILE_IMAGE = 4378000.0  # (475.0 MB read + 926.1 MB written) / 320 images, 8 k_tile launches of one ncu --set full capture
TRAFFIC_TILE_SOURCE = "profiles/r2_final_ncu_full_k_tile.csv (dram__bytes_read.sum + dram__bytes_write.sum over the 8 level launches)"
B_FAST_IMAGE = 1444097           # FAST stage: every pyramid pixel read once (sum of the 8 level sizes)
# dram__bytes_read.sum + dram__bytes_write.sum of one 320-image k_fast_cells launch (ncu --set full,
# profiles/r1_ncu_full_k_fast_cells_v15.csv: 408.31 MB + 46.83 MB), per image
TRAFFIC_FAST_IMAGE = (408313600 + 46831616) / 320.0
# SURVEY.md §8(d), LocalBA 50 keyframes / 5000 points / 30 k stereo edges, per LM trial: buildSystem 6.1 MB + Schur 5.4 MB +
# back-substitution 4.3 MB ~= 16 MB, ~= 64 MFLOP (FP64)
BA_BYTES_TRIAL = 16.0e6
BA_FLOP_TRIAL = 64.0e6
FP64_NOMINAL_TFLOPS = 37.0       # B200 data sheet (HGX B200: 296 TFLOP/s FP64 over 8 GPUs); not in MEASURED_PEAKS.json
# dram__bytes_read.sum + dram__bytes_write.sum of one k_local_ba launch (32 windows, 480 LM trials; ncu --set full,
# profiles/r1_ncu_full_k_local_ba_v19.csv: 1.57 GB + 1.48 GB; it was 12.58 + 4.62 GB before W_e was recomputed, v18), per LM trial
BA_TRAFFIC_TRIAL = (1.569211e9 + 1.483268e9) / 480.0
BA_EVERY = 5


def make_stream_images(n_distinct, seed0=0):
    """uint8 [2, n_distinct, h, w]: left and right images of n_distinct DIFFERENT stereo frames (seed = frame index),
    generated on a few host processes (38 ms per pair on one core)."""
    from synth import synth_stereo_kitti
    out = np.empty((2, n_distinct, H_IMG, W_IMG), np.uint8)
    seeds = [seed0 + i for i in range(n_distinct)]
    workers = max(1, min(16, (os.cpu_count() or 1) // 2, n_distinct // 8))
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            for i, (l, r) in enumerate(pool.imap(synth_stereo_kitti, seeds, chunksize=4)):
                out[0, i], out[1, i] = l, r
    else:
        for i, sd in enumerate(seeds):
            out[0, i], out[1, i] = synth_stereo_kitti(sd)
    return out


def make_images(n_distinct, total, seed0=0):
    """[2*total, h, w] uint8: L_0..L_{total-1}, R_0..R_{total-1}; n_distinct stereo pairs tiled."""
    base = make_stream_images(n_distinct, seed0)
    out = np.empty((2 * total, H_IMG, W_IMG), np.uint8)
    for i in range(total):
        out[i] = base[0, i % n_distinct]
        out[total + i] = base[1, i % n_distinct]
    return out


def ba_window():
    from synth import synth_local_ba
    return synth_local_ba(n_kf=50, n_fixed=10, n_mp=5000, obs_per_mp=6, seed=42)


def ba_windows(n, seed0=0):
    """n different LocalBA windows around the BASELINE.json shape (50 KF / 5000 MP / 30 k edges): 30-60 keyframes, 3000-6000
    points, 4-7 observations per point, 0-8 % gross outliers, a fifth of them with monocular observations — so the LM traces
    (accepted / rejected trials, early terminations) differ between the windows of one launch."""
    from synth import synth_local_ba_fast
    out = []
    for i in range(n):
        rng = np.random.RandomState(7000 + seed0 + i)
        nkf = int(rng.randint(30, 61))
        out.append(synth_local_ba_fast(n_kf=nkf, n_fixed=int(rng.randint(4, max(5, nkf // 4))),
                                       n_mp=int(rng.randint(3000, 6001)), obs_per_mp=int(rng.randint(4, 8)),
                                       seed=9000 + seed0 + i, mono_frac=0.2 if i % 5 == 4 else 0.0,
                                       outlier_frac=float(rng.uniform(0.0, 0.08))))
    return out


def pose_problems(n, seed0=0):
    """n different PoseOptimization problems (2000 features per frame, 50-70 % with a map point, 5-20 % gross outliers)."""
    from synth import synth_pose_problem
    out = []
    for i in range(n):
        rng = np.random.RandomState(11000 + seed0 + i)
        out.append(synth_pose_problem(n=2000, seed=13000 + seed0 + i, mp_frac=float(rng.uniform(0.5, 0.7)),
                                      mono_frac=0.2, outlier_frac=float(rng.uniform(0.05, 0.2)),
                                      pert_t=float(rng.uniform(0.02, 0.15)), pert_deg=float(rng.uniform(0.2, 1.5))))
    return out


def effective_cpus():
    """Host CPUs this process can actually use: the logical CPU count capped by the cgroup CPU quota (the GPU boxes expose
    128 logical CPUs under a 16-CPU quota: 128 busy threads there are time-sliced onto 16 CPUs' worth of time)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(per)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / per
        except Exception:
            pass
    eff = n if quota is None else max(1, min(n, int(quota + 0.999)))
    return eff, n, quota


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


class CpuArm:
    """The reference's own CPU implementation of the step on the host cores (oracle/ref_stream2_glue.cpp ->
    oracle/_ref/libref_stream2.so, built where /root/reference is mounted; the files travel to the GPU box):
    reference Frame constructor (two ORBextractor threads + ComputeStereoMatches), reference ORBmatcher::SearchByBoW,
    Optimizer::PoseOptimization and Optimizer::LocalBundleAdjustment.  For the two optimizers both builds are timed on one
    problem — the reference's Optimizer.cc + g2o on the Eigen stand-in (libref_optimizer.so) and the oracle's C++ port —
    and the FASTER one runs in the arm (the stand-in Eigen is not vectorised: the port usually wins), so the CPU figure is
    not held down by test scaffolding.  The cv stand-in's resize / FAST / blur are scalar C++ (OpenCV's are SIMD): the
    `cv2_extract_ratio` field measures that gap on this box."""

    def __init__(self):
        import oracle_binding
        self.ob = oracle_binding
        vp = ctypes.c_void_p
        lib2 = os.path.join(ROOT, "oracle", "_ref", "libref_stream2.so")
        if not os.path.exists(lib2):
            raise SystemExit("bench.py: oracle/_ref/libref_stream2.so is missing (python __graft_entry__.py builds it where "
                             "/root/reference is mounted)")
        self.R = ctypes.CDLL(lib2)
        self.R.ref_stream2_step.restype = ctypes.c_double
        self.R.ref_stream2_step.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, vp,
                                            vp, ctypes.c_int, vp, vp]
        self.O = oracle_binding.load().L
        opt = os.path.join(ROOT, "oracle", "_ref", "libref_optimizer.so")
        self.G = ctypes.CDLL(opt) if os.path.exists(opt) else None
        self.keep = []

    class Cfg(ctypes.Structure):
        _fields_ = [("nfeatures", ctypes.c_int), ("nlevels", ctypes.c_int), ("iniTh", ctypes.c_int), ("minTh", ctypes.c_int),
                    ("scaleFactor", ctypes.c_float), ("fx", ctypes.c_float), ("fy", ctypes.c_float), ("cx", ctypes.c_float),
                    ("cy", ctypes.c_float), ("bf", ctypes.c_float), ("thDepth", ctypes.c_float), ("project", ctypes.c_int),
                    ("projTh", ctypes.c_float), ("Tcl", ctypes.c_float * 12)]

    def _ba_array(self, windows):
        ob = self.ob
        arr = (ob.BaProblem * len(windows))()
        for i, d in enumerate(windows):
            k = dict(Tcw=np.ascontiguousarray(d["Tcw"], np.float32), fixed=np.ascontiguousarray(d["fixed"], np.uint8),
                     points=np.ascontiguousarray(d["points"], np.float32), edges=np.ascontiguousarray(d["edges"]))
            self.keep.append(k)
            arr[i] = ob.BaProblem(d["n_kf"], d["n_local"], k["Tcw"].ctypes.data, k["fixed"].ctypes.data, len(k["points"]),
                                  k["points"].ctypes.data, len(k["edges"]), k["edges"].ctypes.data, d["fx"], d["fy"], d["cx"],
                                  d["cy"], d["bf"], 5, 10)
        return arr

    def _pose_array(self, poses):
        ob = self.ob
        arr = (ob.PoseProblem * len(poses))()
        for i, d in enumerate(poses):
            a = ob.pose_problem_arrays(d)
            self.keep.append(a)
            arr[i] = ob.PoseProblem(a["Tcw"].ctypes.data, len(a["has_mp"]), a["has_mp"].ctypes.data, a["Xw"].ctypes.data,
                                    a["kpx"].ctypes.data, a["kpy"].ctypes.data, a["uright"].ctypes.data,
                                    a["inv_sigma2"].ctypes.data, d["fx"], d["fy"], d["cx"], d["cy"], d["bf"])
        return arr

    def pick_solvers(self, window, pose):
        """Single-thread time of one LocalBA window / one PoseOptimization with both builds; the faster one is used."""
        ob = self.ob
        info = {}
        t = time.perf_counter(); ob.call_local_ba(self.O.orc_local_ba, window); info["local_ba_port_ms"] = (time.perf_counter() - t) * 1e3
        t = time.perf_counter(); ob.call_pose_optimization(self.O.orc_pose_optimization, pose); info["pose_port_ms"] = (time.perf_counter() - t) * 1e3
        self.ba_fn, self.pose_fn = self.O.orc_local_ba, self.O.orc_pose_optimization
        info["local_ba_impl"] = info["pose_impl"] = "oracle port"
        if self.G is not None:
            t = time.perf_counter(); ob.call_local_ba(self.G.ref_local_ba, window); info["local_ba_ref_g2o_ms"] = (time.perf_counter() - t) * 1e3
            t = time.perf_counter(); ob.call_pose_optimization(self.G.ref_pose_optimization, pose); info["pose_ref_g2o_ms"] = (time.perf_counter() - t) * 1e3
            if info["local_ba_ref_g2o_ms"] < info["local_ba_port_ms"]:
                self.ba_fn, info["local_ba_impl"] = self.G.ref_local_ba, "reference Optimizer.cc + g2o (Eigen stand-in)"
            if info["pose_ref_g2o_ms"] < info["pose_port_ms"]:
                self.pose_fn, info["pose_impl"] = self.G.ref_pose_optimization, "reference Optimizer.cc + g2o (Eigen stand-in)"
        return info

    def prepare(self, windows, poses):
        self.nBa, self.nPose = len(windows), len(poses)
        self.ba_arr = self._ba_array(windows) if windows else None
        self.pose_arr = self._pose_array(poses) if poses else None
        self.cfg = CpuArm.Cfg(NFEAT, 8, 20, 7, 1.2, FX, FY, CX, CY, BF, 35.0, 1, PROJ_TH,
                              (ctypes.c_float * 12)(*[float(v) for v in STREAM_MOTION.reshape(12)]))

    def step(self, imgs_lr, threads):
        """imgs_lr: uint8 [2, S, h, w] contiguous.  Returns (wall seconds, stats[12])."""
        S = imgs_lr.shape[1]
        stats = (ctypes.c_double * 12)()
        vp = ctypes.c_void_p
        wall = self.R.ref_stream2_step(ctypes.byref(self.cfg), imgs_lr.ctypes.data_as(vp), S, W_IMG, H_IMG, threads,
                                       ctypes.cast(self.ba_arr, vp) if self.ba_arr is not None else None, self.nBa,
                                       ctypes.cast(self.ba_fn, vp), ctypes.cast(self.pose_arr, vp) if self.pose_arr is not None else None,
                                       self.nPose, ctypes.cast(self.pose_fn, vp), stats)
        return wall, list(stats)


def cv2_extract_ratio():
    """How much faster OpenCV's own (SIMD) resize / FAST / GaussianBlur are than the scalar cv stand-in the reference's
    ORBextractor.cc is compiled against here: single-thread time of those three stages on one 1241x376 image, both ways
    (Python cv2 is the OpenCV build available in this image).  None if cv2 is not importable."""
    try:
        import cv2
        import oracle_binding
        from synth import synth_image
    except Exception:
        return None
    cv2.setNumThreads(1)
    o = oracle_binding.load()
    img = synth_image(W_IMG, H_IMG, 3)
    lv = [img]
    for l in range(1, 8):
        s = 1.2 ** l
        lv.append(cv2.resize(lv[-1], (int(round(W_IMG / s)), int(round(H_IMG / s))), interpolation=cv2.INTER_LINEAR))
    fd20 = cv2.FastFeatureDetector_create(20, True)

    def run_cv2():
        cur = img
        for l in range(8):
            if l:
                cur = cv2.resize(cur, (lv[l].shape[1], lv[l].shape[0]), interpolation=cv2.INTER_LINEAR)
            fd20.detect(cur[16:-16, 16:-16])
            cv2.GaussianBlur(cur, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)

    def run_stand_in():
        cur = img
        for l in range(8):
            if l:
                cur = o.resize(cur, lv[l].shape[1], lv[l].shape[0])
            o.fast(np.ascontiguousarray(cur[16:-16, 16:-16]), 20)
            o.blur(cur)
    out = {}
    for name, fn in (("cv2_ms", run_cv2), ("stand_in_ms", run_stand_in)):
        fn()
        t = time.perf_counter()
        for _ in range(3):
            fn()
        out[name] = (time.perf_counter() - t) / 3 * 1e3
    out["ratio"] = out["stand_in_ms"] / max(out["cv2_ms"], 1e-9)
    out["note"] = ("whole-level FAST instead of the reference's per-cell calls on both sides; pyramid + FAST(20) + blur of one "
                   "1241x376 image, one thread")
    return out


def cpu_arm_measure(S, steps, warmup, seed0=0, with_cv2=True):
    """`steps` timed steps of S stereo frames each on all host threads; returns the fields of the reference line."""
    threads, logical, quota = effective_cpus()
    arm = CpuArm()
    windows = ba_windows((S + BA_EVERY - 1) // BA_EVERY, seed0)
    poses = pose_problems(S, seed0)
    info = arm.pick_solvers(windows[0], poses[0])
    arm.prepare(windows, poses)
    imgs = np.ascontiguousarray(make_stream_images(S, seed0))
    for _ in range(warmup):
        arm.step(imgs, threads)
    tot, busy = 0.0, np.zeros(5)
    last = None
    for _ in range(steps):
        wall, st = arm.step(imgs, threads)
        tot += wall
        busy += np.array(st[:4] + [st[8]])
        last = st
    fps = S * steps / tot
    stage = {"frame_extract_stereo_ms_per_frame": 1e3 * busy[0] / (S * steps), "search_by_bow_ms_per_frame": 1e3 * busy[1] / (S * steps),
             "pose_optimization_ms_per_frame": 1e3 * busy[2] / (S * steps),
             "search_by_projection_ms_per_frame": 1e3 * busy[4] / (S * steps),
             "local_ba_ms_per_window": 1e3 * busy[3] / max(1, len(windows) * steps)}
    cv2r = cv2_extract_ratio() if with_cv2 else None
    est = None
    if cv2r and cv2r.get("ratio", 0) > 1:
        # what the same arm would deliver if the three OpenCV primitives the reference calls (resize, FAST, GaussianBlur) ran
        # at the speed of OpenCV's own SIMD code instead of the scalar stand-in: they are 2 images x stand_in_ms of the
        # per-frame extraction time; everything else (quadtree, orientation, rBRIEF, matching, solvers) is the reference's
        # own code either way
        per_frame_ms = 1e3 * busy.sum() / (S * steps)
        saved = min(2.0 * cv2r["stand_in_ms"] * (1.0 - 1.0 / cv2r["ratio"]), 0.9 * stage["frame_extract_stereo_ms_per_frame"])
        est = {"fps": fps * per_frame_ms / max(per_frame_ms - saved, 1e-9), "busy_ms_per_frame": per_frame_ms,
               "ms_per_frame_saved": saved,
               "how": "busy ms per frame minus 2 images x stand_in_ms x (1 - 1/ratio), same utilisation"}
    return {"fps": fps, "estimate_with_opencv_simd_primitives": est, "threads": threads, "logical_cpus": logical, "cgroup_cpu_quota": quota, "seconds": tot, "S": S, "tasks_per_step": 3 * S + len(windows) + len(poses),
            "utilisation": float(busy.sum() / (threads * tot)), "stage_ms_single_thread": stage, "solvers": info,
            "keypoints_per_image": last[4], "stereo_matches_per_frame": last[5], "bow_matches_per_frame": last[6], "projection_matches_per_frame": last[9],
            "cv2": cv2r}


def cpu_sample_text(m):
    return ("%d stereo frames per step (%d tasks on %d host threads = the usable CPUs [%d logical, cgroup quota %s], %.1f s timed, "
            "utilisation %.0f %%): reference's own Frame "
            "constructor (ORBextractor.cc x2 threads + ComputeStereoMatches), ORBmatcher::SearchByBoW, ORBmatcher::SearchByProjection"
            "(CurrentFrame, LastFrame), PoseOptimization [%s], "
            "LocalBundleAdjustment [%s]; OpenCV primitives = scalar stand-in (cv2 is %sx faster on pyramid+FAST+blur%s)"
            % (m["S"], m["tasks_per_step"], m["threads"], m["logical_cpus"], m["cgroup_cpu_quota"], m["seconds"],
               100 * m["utilisation"], m["solvers"]["pose_impl"],
               m["solvers"]["local_ba_impl"], ("%.1f" % m["cv2"]["ratio"]) if m.get("cv2") else "n/a ",
               ("; with those at OpenCV speed the arm would reach about %.0f frames/s" % m["estimate_with_opencv_simd_primitives"]["fps"])
               if m.get("estimate_with_opencv_simd_primitives") else ""))


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the workload on all host cores (see CpuArm)."""
    if rank != 0:
        return
    S = args.ref_frames if args.ref_frames > 0 else args.frames
    m = cpu_arm_measure(S, args.steps, args.warmup)
    fps = m["fps"]
    line = {
        "impl": "reference", "metric": "stereo_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * m["seconds"] / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (extract, match), f64 (LocalBA)", "data": "synthetic",
        "config": workload_config(S, 1),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": m["threads"], "kind": "reference", "sample": cpu_sample_text(m),
                         "detail": m},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(frames_per_gpu, world):
    return {"workload": "batched KITTI-shape stereo stream 1241x376, 2000 feat/img, 8 levels, FAST 20/7, distinct frames: extract "
                        "L+R, ComputeStereoMatches, temporal SearchByBoW 2000x2000 (one vocabulary node) and SearchByProjection(CurrentFrame, "
                        "LastFrame) of the last frame's stereo points (motion model, th 7), PoseOptimization per "
                        "frame, LocalBA every 5th frame (32 different windows per 160 frames, 30-60 KF / 3000-6000 MP / ~30k edges)",
            "frames_per_step_per_gpu": frames_per_gpu, "images_per_step_per_gpu": 2 * frames_per_gpu,
            "parallelism": "frames sharded x%d, NCCL all-gather of the shard-boundary left-image feature records" % world,
            "l2": "inputs per step (%.0f MB of images per GPU) exceed the 126 MB L2"
                  % (2 * frames_per_gpu * W_IMG * H_IMG / 1e6)}


def run_b200(args, rank, local_rank, world):
    import torch
    pkg = importlib.import_module("self_commit_orb-slam2_b200")
    stream_mod = importlib.import_module("self_commit_orb-slam2_b200.stream")
    if pkg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    F, D = args.frames, max(args.frames, args.distinct)
    n_ba = (F + BA_EVERY - 1) // BA_EVERY
    windows = ba_windows(n_ba, seed0=1000 * rank)
    poses = pose_problems(F, seed0=1000 * rank)
    if world > 1:  # several ranks share the host cores: split them for the LocalBA window preparation threads
        os.environ.setdefault("B2S_BA_HOST_THREADS", str(max(2, min(16, effective_cpus()[0] // world))))
    ss_kw = dict(ba_problems=windows, pose_problems=poses, stereo=True, ba_every=BA_EVERY, device=local_rank, rank=rank,
                 world=world, ba_depth=args.ba_depth, exchange=args.exchange, bf=BF, project=True, intrinsics=(FX, FY, CX, CY),
                 motion=STREAM_MOTION, ba_sms=args.ba_sms)
    ss = stream_mod.StereoStream(F, W_IMG, H_IMG, NFEAT, **ss_kw)
    imgs = make_stream_images(D, seed0=100000 * rank)  # [2, D, h, w], frame index = seed
    pinned = torch.from_numpy(imgs).pin_memory()
    d_all = pinned.cuda(non_blocking=True)
    torch.cuda.synchronize()
    step_no = [0]

    # `--dev-lanes` device-resident pipelines (own handles and streams) take the steps in turn: the low-occupancy kernels of one
    # step (greedy resolvers, quadtree, PoseOptimization) then overlap the wide kernels of the next
    dev_lanes = [ss]
    for _ in range(max(1, args.dev_lanes) - 1):
        kw = dict(ss_kw)
        kw["ba_depth"] = 1
        dev_lanes.append(stream_mod.StereoStream(F, W_IMG, H_IMG, NFEAT, **kw))

    def dev_step():
        lane = dev_lanes[step_no[0] % len(dev_lanes)]
        lane.load_window(d_all, (step_no[0] * F) % D)
        step_no[0] += 1
        return lane.step_device(pipelined=True)  # solvers of step k overlap extraction / matching of step k+1

    def dev_finish():
        for lane in dev_lanes:
            lane.finish()

    def dev_launches():
        return sum(lane.launch_count() for lane in dev_lanes)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---------------- device-resident throughput (`value`)
    for _ in range(max(args.warmup, len(dev_lanes))):
        dev_step()
    dev_finish()
    ss.ex.check()
    torch.cuda.synchronize()
    barrier()
    L = pkg.lib()
    L.b2s_extractor_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.b2s_extractor_get_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.b2s_extractor_set_timing(ss.ex._h, 1)
    launches0 = dev_launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    ev0.record(ss.stream)
    for _ in range(args.steps):
        dev_step()
    dev_finish()                        # the last solver batches are joined inside the timed region
    ev1.record(ss.stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    # the solver batches run on their own streams and are synchronous on their host thread, so the host wall clock bounds
    # everything
    dev_ms = max(ev0.elapsed_time(ev1), wall * 1e3)
    launches = dev_launches() - launches0
    stage = (ctypes.c_double * 5)()
    calls = ctypes.c_longlong(0)
    L.b2s_extractor_get_timing(ss.ex._h, stage, ctypes.byref(calls))
    L.b2s_extractor_set_timing(ss.ex._h, 0)
    ss.ex.check()
    t_all = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    dev_ms = float(t_all.item())
    value = world * F * args.steps / (dev_ms / 1e3)

    # ---------------- phase breakdown (untimed extra passes, for DESIGN.md / profiles)
    phase = {}
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(2):
        ss.load_window(d_all, 0)
        ss.step_device(run_ba=False)
    torch.cuda.synchronize()
    phase["extract_stereo_match_ms"] = (time.perf_counter() - t1) * 500.0
    L.b2s_extractor_set_timing(ss.ex._h, 1)  # extractor stage timing without the solvers competing for the SMs
    for _ in range(2):
        ss.step_device(run_ba=False)
    torch.cuda.synchronize()
    stage_iso = (ctypes.c_double * 5)()
    calls_iso = ctypes.c_longlong(0)
    L.b2s_extractor_get_timing(ss.ex._h, stage_iso, ctypes.byref(calls_iso))
    L.b2s_extractor_set_timing(ss.ex._h, 0)
    phase["extractor_stage_ms_isolated"] = {k: stage_iso[i] / max(1, calls_iso.value) for i, k in enumerate(STAGE_NAMES)}
    phase["stereo_matches_per_frame"] = float(ss.nstereo.float().mean().item())
    phase["keypoints_per_image"] = float(ss.counts[1:].float().mean().item())
    phase["projection_matches_per_frame"] = float(ss.npmatch.float().mean().item())
    phase["bow_matches_per_frame"] = float(ss.nmatch.float().mean().item())
    try:  # the batched SearchByBoW call alone (rank + K-lists + resolver + cull) on the records of the last step
        _vp = ctypes.c_void_p
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st = ss.stream.cuda_stream

        def bow_call():
            pkg._check(L.b2s_search_by_bow_device(ss.matcher._h, F, _vp(ss.desc.data_ptr()), _vp(ss.node.data_ptr()),
                                                  _vp(ss.valid.data_ptr()), _vp(ss.ang.data_ptr()), _vp(ss.counts.data_ptr()),
                                                  ss.cap, _vp(ss.desc[1:].data_ptr()), _vp(ss.node[1:].data_ptr()), None,
                                                  _vp(ss.ang[1:].data_ptr()), _vp(ss.counts[1:].data_ptr()), ss.cap, 50,
                                                  float(ss.matcher.mfNNratio), 0, 1, _vp(ss.match.data_ptr()),
                                                  _vp(ss.nmatch.data_ptr()), _vp(st)))
        bow_call()
        torch.cuda.synchronize()
        e0.record(ss.stream)
        for _ in range(5):
            bow_call()
        e1.record(ss.stream)
        torch.cuda.synchronize()
        phase["search_by_bow_batch_ms"] = e0.elapsed_time(e1) / 5
        cn = ss.counts.cpu().numpy().astype(np.float64)
        phase["search_by_bow_descriptor_pairs"] = float((cn[:F] * cn[1:1 + F]).sum())
    except Exception as ex:
        phase["search_by_bow_side_measurement_error"] = str(ex)[:200]
    ba_kernel_ms, ba_trials = 0.0, 0
    try:
        t1 = time.perf_counter()
        pr = ss.pose_opt.PoseOptimizationBatch(poses)
        phase["pose_optimization_batch_ms"] = (time.perf_counter() - t1) * 1e3
        phase["pose_optimization_inliers_per_frame"] = float(np.mean([r["n_inliers"] for r in pr]))
        t1 = time.perf_counter()
        outs = ss.opt.LocalBundleAdjustmentBatch(windows)
        phase["local_ba_batch_ms"] = (time.perf_counter() - t1) * 1e3
        phase["local_ba_windows"] = n_ba
        phase["local_ba_trials_per_window"] = [int(o["n_trials"]) for o in outs]
        ba_kernel_ms, ba_trials = ss.opt.last_kernel_ms()
    except Exception as ex:  # never let a side measurement break the bench line
        phase["solver_side_measurement_error"] = str(ex)[:200]
    # the round-1 workload (no stereo matching / PoseOptimization, one LocalBA window x 32) for continuity
    try:
        w1 = ba_window()
        ss1 = ss
        keep = (ss1.windows, ss1.pose_problems, ss1.stereo, ss1.project)
        ss1.windows, ss1.pose_problems, ss1.stereo, ss1.project = [w1], [], False, False
        ss1._prep.clear()
        for _ in range(2):
            ss1.step_device(pipelined=True)
        ss1.finish()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(4):
            ss1.step_device(pipelined=True)
        ss1.finish()
        torch.cuda.synchronize()
        phase["round1_workload_frames_per_s"] = 4 * F / (time.perf_counter() - t1)
        ss1.windows, ss1.pose_problems, ss1.stereo, ss1.project = keep
        ss1._prep.clear()
    except Exception as ex:
        phase["round1_workload_error"] = str(ex)[:200]

    # ---------------- end to end through the host-buffer C ABI (`e2e`)
    e2e_steps = args.steps if args.e2e_steps <= 0 else max(1, min(args.steps, args.e2e_steps))
    base_ptr, img_bytes = pinned.data_ptr(), W_IMG * H_IMG  # e2e inputs come from pinned host memory

    # `--e2e-lanes` host-API pipelines (own handles, own pinned result buffers) take the steps in turn, each on its own host
    # thread, so that the transfers of one step overlap the kernels of the next; every step still uploads its 2F images from
    # pinned host memory and downloads all its results inside the timed region
    lanes = list(dev_lanes[:max(1, args.e2e_lanes)])
    for _ in range(max(1, args.e2e_lanes) - len(lanes)):
        kw = dict(ss_kw)
        kw["ba_depth"] = 1
        lanes.append(stream_mod.StereoStream(F, W_IMG, H_IMG, NFEAT, **kw))

    def host_step(lane, k):
        idx = [(k * F + i) % D for i in range(F)]
        ptrs = [base_ptr + i * img_bytes for i in idx] + [base_ptr + (D + i) * img_bytes for i in idx]
        return lane.step_host(None, pipelined=True, img_ptrs=ptrs)

    def lane_loop(j, first, last):
        for k in range(first + j, last, len(lanes)):
            host_step(lanes[j], k)
        lanes[j].finish()
    for j, lane in enumerate(lanes):  # warm
        host_step(lane, j)
        lane.finish()
    torch.cuda.synchronize()
    barrier()
    import threading
    t0 = time.perf_counter()
    th = [threading.Thread(target=lane_loop, args=(j, 1, 1 + e2e_steps)) for j in range(1, len(lanes))]
    for t in th:
        t.start()
    lane_loop(0, 1, 1 + e2e_steps)
    for t in th:
        t.join()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * F * e2e_steps / float(t_e.item())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks, peak_src = load_peaks()
    micro = measure_device_peaks(pkg)
    # ---- roofline of the dominant front-end kernel: the fused tile kernel (8 launches, one per level)
    tile_ms = stage[1] / max(1, calls.value)
    tile_iso = phase["extractor_stage_ms_isolated"].get(STAGE_NAMES[1])
    images_per_launch = 2 * F
    achieved = B_TILE_IMAGE * images_per_launch / (tile_ms * 1e-3) / 1e9 if tile_ms > 0 else 0.0
    ex_ms = sum(stage) / max(1, calls.value)
    roof = {"kernel": "k_tile x 8 levels + k_cells + k_fast_cells_list (TMA-staged tile: FAST strength + NMS bitmap + Q8 blur + "
                      "next pyramid level; per-cell threshold rules)", "bound": "hbm", "achieved": achieved,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": TRAFFIC_TILE_IMAGE * images_per_launch,
            "traffic_source": TRAFFIC_TILE_SOURCE,
            "frac_isolated": (B_TILE_IMAGE * images_per_launch / (tile_iso * 1e-3) / 1e9 / peaks["hbm_gbs"]) if tile_iso else None,
            "frac_note": "launch_ms is measured inside the timed steps, where LocalBA holds %s of the 148 SMs exclusively; "
                         "launch_ms_isolated / frac_isolated are the same launches without the solvers running" % (args.ba_sms if args.ba_sms > 0 else "all"),
            "traffic_frac": (TRAFFIC_TILE_IMAGE * images_per_launch / (tile_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if tile_ms > 0 else None,
            "peak_source": peak_src, "launch_ms": tile_ms, "launch_ms_isolated": tile_iso,
            "algorithmic_bytes_per_launch": B_TILE_IMAGE * images_per_launch,
            "algorithmic_bytes_note": "SURVEY 8d per image: pyramid R+W 2,385,248 + FAST read 1,444,097 + blur R+W 2,888,194 B",
            "issue_bound": {"note": "the tile kernel is bound by the ALU pipe (packed u16x2 min/max of FAST), not by HBM",
                            "alu_pipe_peak_gwarpinstr_s": micro.get("alu_vimnmx3_gwarp_s")},
            "extractor_all_stages": {"ms_per_launch_set": ex_ms,
                                     "achieved_GBps": B_STAGE_IMAGE * images_per_launch / (ex_ms * 1e-3) / 1e9 if ex_ms > 0 else 0.0,
                                     "frac": (B_STAGE_IMAGE * images_per_launch / (ex_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]) if ex_ms > 0 else 0.0,
                                     "stage_ms": {k: stage[i] / max(1, calls.value) for i, k in enumerate(STAGE_NAMES)}}}
    line = {
        "metric": "stereo_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32 (extract, match), f64 (LocalBA)", "data": "synthetic",
        "config": workload_config(F, world),
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": ss.h2d_bytes_per_step(),
                "d2h_bytes_per_step": ss.d2h_bytes_per_step(), "steps": e2e_steps, "host_api_lanes": len(lanes)},
        "gpu_launches": int(launches),
        "phase_ms": phase,
        "clocks": clocks,
        "roofline": roof,
        "device_peaks_measured": micro,
    }
    line["config"]["distinct_frames_resident"] = D
    line["config"]["pipelines"] = ("%d device-resident lanes (value) / %d host-API lanes (e2e) take the steps in turn; LocalBA batches on a "
                                   "budget of %s SMs" % (len(dev_lanes), len(lanes), args.ba_sms if args.ba_sms > 0 else "all"))
    if ba_kernel_ms > 0:
        ba_gbs = BA_BYTES_TRIAL * ba_trials / (ba_kernel_ms * 1e-3) / 1e9
        fp64_peak = micro.get("fp64_dfma_tflops") or FP64_NOMINAL_TFLOPS
        line["roofline_local_ba"] = {
            "kernel": "k_local_ba (persistent LM loop: %d different windows on %s SMs dealt by estimated cost, one launch per "
                      "LocalBA batch)" % (n_ba, args.ba_sms if args.ba_sms > 0 else "all"),
            "bound": "hbm", "achieved": ba_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ba_gbs / peaks["hbm_gbs"],
            "traffic": 5.21e6 * ba_trials, "traffic_source": "profiles/r2_final_ncu_full_k_local_ba.csv: 1.268 GB read + 1.242 GB written per launch of 480 LM trials",
            "traffic_frac": 5.21e6 * ba_trials / (ba_kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "sms_used": (args.ba_sms if args.ba_sms > 0 else 148),
            "frac_of_sm_share": ba_gbs / peaks["hbm_gbs"] / ((args.ba_sms if args.ba_sms > 0 else 148) / 148.0),
            "sm_budget_note": "the batch runs on a budget of SMs (b2s_ba_set_sm_budget) because the step time is the SUM of the SM-time of "
                              "LocalBA and of the front end; on all 148 SMs the same launch takes 7.5 ms (frac 0.156) but the step is slower",
            "tensor_pipe": {"dmma_subpipe_pct": 0.09, "fp64_pipe_pct": 23.7, "source": "profiles/r2_final_ncu_full_k_local_ba.csv "
                            "(sm__inst_executed_pipe_tensor_subpipe_dmma / sm__inst_executed_pipe_fp64, pct of peak sustained active): "
                            "DMMA.8x8x4 carries the trailing update of the reduced-system factorisation only"},
            "launch_ms": ba_kernel_ms, "launch_ms_source": "CUDA events on the solver stream, batch run alone",
            "lm_trials": ba_trials, "algorithmic_bytes_per_launch": BA_BYTES_TRIAL * ba_trials,
            "algorithmic_bytes_note": "SURVEY 8d: 16 MB per LM trial of a 50/5000/30k window (scaled by the launch's trial count)",
            "fp64": {"achieved_tflops": BA_FLOP_TRIAL * ba_trials / (ba_kernel_ms * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                     "peak_source": "measured DFMA micro-benchmark in this run" if micro.get("fp64_dfma_tflops") else "nominal",
                     "frac": BA_FLOP_TRIAL * ba_trials / (ba_kernel_ms * 1e-3) / 1e12 / fp64_peak},
            "note": "latency-bound (window barriers, dependent FP64 chains at 8 warps/SM): issue slots 25 % busy, FP64 pipe 24 %, "
                    "barrier stall 1.9 warps per issue in the capture (68 SMs)"}
    if phase.get("search_by_bow_batch_ms") and micro.get("imma_hamming_pairs_per_s"):
        pairs = phase["search_by_bow_descriptor_pairs"]
        rate = pairs / (phase["search_by_bow_batch_ms"] * 1e-3)
        line["roofline_matcher"] = {
            "kernel": "batched SearchByBoW call: k_rank_by_key + k_bow_topk_imma (IMMA.16832 K-lists) + k_bow_resolve + k_rot_cull",
            "bound": "tensor", "unit": "descriptor pairs/s", "achieved": rate, "peak": micro["imma_hamming_pairs_per_s"],
            "frac": rate / micro["imma_hamming_pairs_per_s"], "call_ms": phase["search_by_bow_batch_ms"], "pairs_per_call": pairs,
            "peak_source": "measured IMMA.16832.U8.U8 issue rate in this run x 16 pairs per instruction (8 IMMA per 16 x 8 tile of "
                           "256-bit distances)",
            "note": "the K-list kernel is about 40 % of the call (profiles/README.md); the rest are the rank and the greedy resolver"}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(F)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def measure_device_peaks(pkg):
    """Issue-rate / FP64 / POPC micro-benchmarks of THIS device in THIS run (b2s_measure_peaks): the bounds the integer
    front end, the FP64 solver and the Hamming matcher actually run against (MEASURED_PEAKS.json has HBM and bf16 only)."""
    try:
        L = pkg.lib()
        out = (ctypes.c_double * 8)()
        L.b2s_measure_peaks.argtypes = [ctypes.c_int, ctypes.c_void_p]
        rc = L.b2s_measure_peaks(0, out)
        if rc != 0:
            return {"error": "b2s_measure_peaks rc=%d" % rc}
        return {"alu_vimnmx3_gwarp_s": out[0], "fma_imad_gwarp_s": out[1], "fp64_dfma_tflops": out[2], "popc_gwarp_s": out[3],
                "dual_issue_alu_fma_gwarp_s": out[4], "imma_16832_gwarp_s": out[5],
                "imma_hamming_pairs_per_s": out[5] * 1e9 * 16,  # 8 IMMA per 16 x 8 tile of 256-bit distances
                "how": "b2s_measure_peaks: dependent-free unrolled chains, 148 x 8 CTAs x 256 threads, CUDA events, best of 3"}
    except Exception as ex:
        return {"error": str(ex)[:200]}


def cpu_baseline(F):
    """The CPU arm on this box's host cores, bounded sample: one warm-up + two timed steps of the SAME workload (F frames per
    step); see CpuArm."""
    m = cpu_arm_measure(F, steps=2, warmup=1, with_cv2=True)
    return {"value": m["fps"], "unit": "frames/s", "cores": m["threads"], "kind": "reference", "sample": cpu_sample_text(m),
            "detail": m}


def main():
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=160, help="stereo frames per step per GPU (160 -> 149 MB of images)")
    ap.add_argument("--ref-frames", type=int, default=0, help="stereo frames per step of the CPU reference arm (0: --frames)")
    ap.add_argument("--distinct", type=int, default=512, help="different stereo frames resident per GPU (seed = frame index)")
    ap.add_argument("--dev-lanes", type=int, default=int(os.environ.get("B2S_DEV_LANES", "2")),
                    help="device-resident pipelines that take the timed steps in turn")
    ap.add_argument("--e2e-lanes", type=int, default=int(os.environ.get("B2S_E2E_LANES", "2")),
                    help="host-API pipelines that take the e2e steps in turn (double buffering of transfers and kernels)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end loop (0: the same K as --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="boundary", choices=["boundary", "all"],
                    help="NCCL all-gather of the shard-boundary left-image record only, or of every left-image record")
    ap.add_argument("--ba-sms", type=int, default=int(os.environ.get("B2S_BENCH_BA_SMS", "64")),
                    help="SMs one LocalBA batch may occupy (b2s_ba_set_sm_budget; 0 = all)")
    ap.add_argument("--ba-depth", type=int, default=int(os.environ.get("B2S_BA_DEPTH", "2")),
                    help="LocalBA solver handles used round-robin by the pipelined stream (host work of batch i+1 overlaps the kernel of batch i)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
