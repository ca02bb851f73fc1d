#!/usr/bin/env python3
"""bench.py — stereo frames/s of the hot path (extract + match + LocalBA) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                 # this framework (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus 1 --steps K --warmup W  # the reference algorithm on the host cores (oracle)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one batch of F synthetic KITTI-shaped stereo frames (1241x376, 2000 features, 8 levels) per GPU:
extraction of the 2F images, temporal SearchByBoW (2000x2000 brute force, one vocabulary node) of every left image
against its predecessor, one LocalBA (50 KF / 5000 MP / 30k edges) per 5 frames; with N > 1 the left-image feature
records are all-gathered over NCCL.  Prints ONE JSON line (see DESIGN.md "Measurement").
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
# algorithmic bytes per 1241x376 image
B_STAGE_IMAGE = 9359539          # all extractor stages
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


def make_images(n_distinct, total, seed0=0):
    """[2*total, h, w] uint8: L_0..L_{total-1}, R_0..R_{total-1}; n_distinct stereo pairs tiled."""
    from synth import synth_stereo
    base = [synth_stereo(W_IMG, H_IMG, seed0 + i) for i in range(n_distinct)]
    out = np.empty((2 * total, H_IMG, W_IMG), np.uint8)
    for i in range(total):
        l, r = base[i % n_distinct]
        out[i] = l
        out[total + i] = r
    return out


def ba_window():
    from synth import synth_local_ba
    return synth_local_ba(n_kf=50, n_fixed=10, n_mp=5000, obs_per_mp=6, seed=42)


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


def cpu_stream_step_fn():
    """The CPU arm's step function and what it is made of.  Preferred: oracle/_ref/libref_stream.so — the reference's OWN
    src/ORBextractor.cc and src/ORBmatcher.cc compiled in place (oracle/Makefile `ref`, built where /root/reference is
    mounted; the file travels to the GPU box) with the oracle port of LocalBA (src/Optimizer.cc needs g2o + Eigen, absent
    here).  Otherwise the oracle port of all three stages (oracle/orb_misc.cpp orc_stream_step)."""
    import oracle_binding
    vp = ctypes.c_void_p
    argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int,
                ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp]
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libref_stream.so")
    if os.path.exists(ref_lib) and not os.environ.get("B2S_BENCH_CPU_PORT"):
        R = ctypes.CDLL(ref_lib)
        fn = R.ref_stream_step
        kind = "reference"
        what = ("reference's own ORBextractor.cc + ORBmatcher.cc compiled in place (oracle/_ref) + oracle port of LocalBA "
                "(Optimizer.cc needs Eigen)")
    else:
        fn = oracle_binding.load().L.orc_stream_step
        kind = "port"
        what = "oracle port (extract, match, LocalBA)"
    fn.restype = ctypes.c_double
    fn.argtypes = argtypes
    ba = ba_window()
    keep = dict(Tcw=np.ascontiguousarray(ba["Tcw"], np.float32), fixed=np.ascontiguousarray(ba["fixed"], np.uint8),
                points=np.ascontiguousarray(ba["points"], np.float32), edges=np.ascontiguousarray(ba["edges"]))
    prob = oracle_binding.BaProblem(ba["n_kf"], ba["n_local"], keep["Tcw"].ctypes.data, keep["fixed"].ctypes.data,
                                    len(keep["points"]), keep["points"].ctypes.data, len(keep["edges"]),
                                    keep["edges"].ctypes.data, ba["fx"], ba["fy"], ba["cx"], ba["cy"], ba["bf"], 5, 10)

    def step(imgs, S, threads):
        assert keep is not None  # (the problem arrays must outlive the call)
        return fn(NFEAT, 1.2, 8, 20, 7, imgs.ctypes.data_as(vp), S, W_IMG, H_IMG, ctypes.byref(prob), BA_EVERY, threads, None)
    return step, kind, what


def run_reference(args, rank, world):
    """The reference algorithm's CPU path on all host cores (see cpu_stream_step_fn)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    S = args.ref_frames
    imgs = make_images(min(S, 8), S)
    step, kind, what = cpu_stream_step_fn()
    for _ in range(args.warmup):
        step(imgs, S, threads)
    t = 0.0
    for _ in range(args.steps):
        t += step(imgs, S, threads)
    fps = S * args.steps / t
    line = {
        "impl": "reference", "metric": "stereo_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (extract, match), f64 (LocalBA)", "data": "synthetic",
        "config": workload_config(S, 1),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": "%d stereo frames per step (bounded sample of the workload) on %d host threads: %s"
                                   % (S, threads, what)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(frames_per_gpu, world):
    return {"workload": "batched KITTI-shape stereo stream 1241x376, 2000 feat/img, 8 levels, FAST 20/7: extract L+R, "
                        "temporal SearchByBoW 2000x2000 (one vocabulary node), LocalBA 50KF/5000MP/30k edges every 5th frame",
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
    F = args.frames
    ba = ba_window()
    if world > 1:  # several ranks share the host cores: split them for the LocalBA window preparation threads
        os.environ.setdefault("B2S_BA_HOST_THREADS", str(max(2, min(16, (os.cpu_count() or 16) // world))))
    ss = stream_mod.StereoStream(F, W_IMG, H_IMG, NFEAT, ba_problem=ba, ba_every=BA_EVERY, device=local_rank, rank=rank,
                                 world=world, ba_depth=args.ba_depth, exchange=args.exchange)
    imgs = make_images(min(F, 16), F, seed0=1000 * rank)
    pinned = torch.from_numpy(imgs).pin_memory()
    ss.upload(pinned)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---------------- device-resident throughput (`value`)
    for _ in range(args.warmup):
        ss.step_device(pipelined=True)
    ss.finish()
    ss.ex.check()
    torch.cuda.synchronize()
    barrier()
    L = pkg.lib()
    L.b2s_extractor_set_timing.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.b2s_extractor_get_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.b2s_extractor_set_timing(ss.ex._h, 1)
    launches0 = ss.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    ev0.record(ss.stream)
    for _ in range(args.steps):
        ss.step_device(pipelined=True)  # LocalBA of step k overlaps extraction/matching of step k+1
    ss.finish()                         # ... and the last batch is joined inside the timed region
    ev1.record(ss.stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    # the LocalBA batch runs on its own stream and is synchronous on the host, so the host wall clock bounds everything
    dev_ms = max(ev0.elapsed_time(ev1), wall * 1e3)
    launches = ss.launch_count() - launches0
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

    # ---------------- phase breakdown (untimed extra passes, for DESIGN.md / profiles): extraction+matching alone, LocalBA alone
    phase = {}
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(2):
        ss.step_device(run_ba=False)
    torch.cuda.synchronize()
    phase["extract_match_ms"] = (time.perf_counter() - t1) * 500.0
    # extractor stage timing without LocalBA competing for the SMs
    L.b2s_extractor_set_timing(ss.ex._h, 1)
    for _ in range(2):
        ss.step_device(run_ba=False)
    torch.cuda.synchronize()
    stage_iso = (ctypes.c_double * 5)()
    calls_iso = ctypes.c_longlong(0)
    L.b2s_extractor_get_timing(ss.ex._h, stage_iso, ctypes.byref(calls_iso))
    L.b2s_extractor_set_timing(ss.ex._h, 0)
    phase["extractor_stage_ms_isolated"] = {k: stage_iso[i] / max(1, calls_iso.value) for i, k in enumerate(
        ["resize_chain", "fast_cells", "quadtree", "blur", "orient_describe"])}
    # Frame::ComputeStereoMatches for the F pairs of the step, on the resident pyramids (reported next to
    # the step, not part of the metric): left image i <-> right image F+i
    try:
        ur = torch.zeros((F, ss.cap), dtype=torch.float32, device="cuda")
        dp = torch.zeros((F, ss.cap), dtype=torch.float32, device="cuda")
        nmt = torch.zeros(F, dtype=torch.int32, device="cuda")
        st = ss.stream
        with torch.cuda.stream(st):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for rep in range(2):
                if rep == 1:
                    e0.record(st)
                ss.ex.stereo_match_device(0, F, F, ss.kps[1:].data_ptr(), ss.desc[1:].data_ptr(),
                                          ss.counts[1:].data_ptr(), ss.cap, 386.1448, 0.0, ur.data_ptr(), dp.data_ptr(),
                                          nmt.data_ptr(), stream=st.cuda_stream)
            e1.record(st)
        torch.cuda.synchronize()
        phase["stereo_match_ms_isolated"] = e0.elapsed_time(e1)
        phase["stereo_matches_per_frame"] = float(nmt.float().mean().item())
    except Exception as ex:  # never let the side measurement break the bench line
        phase["stereo_match_error"] = str(ex)[:200]
    # Optimizer::PoseOptimization for the F frames of the step (side measurement through the host-buffer
    # C ABI: packing, H2D, one CTA per frame, D2H)
    try:
        from synth import synth_pose_problem
        pp = [synth_pose_problem(seed=1000 + (i % 8)) for i in range(8)]
        frames_pp = [pp[i % 8] for i in range(F)]
        popt = ss.opt if ss.opt is not None else None
        if popt is not None:
            popt.PoseOptimizationBatch(frames_pp)
            t1 = time.perf_counter()
            pr = popt.PoseOptimizationBatch(frames_pp)
            phase["pose_optimization_batch_ms"] = (time.perf_counter() - t1) * 1e3
            phase["pose_optimization_frames"] = F
            phase["pose_optimization_inliers_per_frame"] = float(np.mean([r["n_inliers"] for r in pr]))
    except Exception as ex:
        phase["pose_optimization_error"] = str(ex)[:200]
    if ss.n_ba:
        t1 = time.perf_counter()
        ss.opt.LocalBundleAdjustmentBatch([ss.ba_problem] * ss.n_ba)
        phase["local_ba_batch_ms"] = (time.perf_counter() - t1) * 1e3
        phase["local_ba_windows"] = ss.n_ba
        try:
            ba_kernel_ms, ba_trials = ss.opt.last_kernel_ms()
        except Exception:
            ba_kernel_ms, ba_trials = 0.0, 0

    # ---------------- end to end through the host-buffer C ABI (`e2e`)
    e2e_steps = args.steps if args.e2e_steps <= 0 else max(1, min(args.steps, args.e2e_steps))
    imgs_pinned = pinned.numpy()  # e2e inputs come from pinned host memory
    ss.step_host(imgs_pinned)  # warm
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        n, nm, ba_out, _ = ss.step_host(imgs_pinned, pipelined=True)
    ss.finish()
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
    roofline_ba = None
    if ss.n_ba and ba_kernel_ms > 0:
        ba_gbs = BA_BYTES_TRIAL * ba_trials / (ba_kernel_ms * 1e-3) / 1e9
        roofline_ba = {"kernel": "k_local_ba (persistent LM loop: %d windows x 4 CTAs, one launch per LocalBA batch)" % ss.n_ba,
                       "bound": "hbm", "achieved": ba_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": ba_gbs / peaks["hbm_gbs"], "traffic": BA_TRAFFIC_TRIAL * ba_trials,
                       "traffic_source": "profiles/r1_ncu_full_k_local_ba_v19.csv (ncu --set full, 32-window launch)",
                       "launch_ms": ba_kernel_ms, "launch_ms_source": "CUDA events on the solver stream, batch run alone",
                       "lm_trials": ba_trials, "algorithmic_bytes_per_launch": BA_BYTES_TRIAL * ba_trials,
                       "fp64": {"achieved_tflops": BA_FLOP_TRIAL * ba_trials / (ba_kernel_ms * 1e-3) / 1e12,
                                "nominal_peak_tflops": FP64_NOMINAL_TFLOPS},
                       "note": "latency-bound (barrier 3.8 + long_scoreboard 2.4 warps per issue at 8 warps/SM, FP64 pipe "
                               "18 %); since W_e is recomputed by its consumers and the error pass before buildSystem is "
                               "reused, DRAM traffic (6.4 MB per LM trial) is below SURVEY's 16 MB estimate, which assumed a "
                               "stored 4.3 MB W array per window (see profiles/README.md)"}

    fast_ms = stage[1] / max(1, calls.value)
    images_per_launch = 2 * F
    achieved = B_FAST_IMAGE * images_per_launch / (fast_ms * 1e-3) / 1e9 if fast_ms > 0 else 0.0
    ex_ms = sum(stage) / max(1, calls.value)
    line = {
        "metric": "stereo_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8/int32 (extract, match), f64 (LocalBA)", "data": "synthetic",
        "config": workload_config(F, world),
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": ss.h2d_bytes_per_step(),
                "d2h_bytes_per_step": ss.d2h_bytes_per_step(), "steps": e2e_steps},
        "gpu_launches": int(launches),
        "phase_ms": phase,
        "clocks": clocks,
        "roofline": {"kernel": "k_fast_cells (per-cell FAST-9/16 score + NMS + dual threshold)", "bound": "hbm",
                     "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": TRAFFIC_FAST_IMAGE * images_per_launch,
                     "traffic_source": "profiles/r1_ncu_full_k_fast_cells_v15.csv (ncu --set full, 320-image launch)",
                     "peak_source": peak_src, "launch_ms": fast_ms,
                     "launch_ms_isolated": phase.get("extractor_stage_ms_isolated", {}).get("fast_cells"),
                     "algorithmic_bytes_per_launch": B_FAST_IMAGE * images_per_launch,
                     "extractor_all_stages": {"ms_per_launch_set": ex_ms,
                                              "achieved_GBps": B_STAGE_IMAGE * images_per_launch / (ex_ms * 1e-3) / 1e9
                                              if ex_ms > 0 else 0.0,
                                              "stage_ms": {"resize_chain": stage[0] / max(1, calls.value),
                                                           "fast_cells": fast_ms,
                                                           "quadtree": stage[2] / max(1, calls.value),
                                                           "blur": stage[3] / max(1, calls.value),
                                                           "orient_describe": stage[4] / max(1, calls.value)}}},
    }
    if roofline_ba is not None:
        line["roofline_local_ba"] = roofline_ba
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline():
    """The CPU arm timed on this box's host cores, bounded sample (about 10-30 core-seconds); see cpu_stream_step_fn."""
    threads = os.cpu_count() or 1
    S = 48
    imgs = make_images(8, S)
    step, kind, what = cpu_stream_step_fn()
    t = step(imgs, S, threads)
    return {"value": S / t, "unit": "frames/s", "cores": threads, "kind": kind,
            "sample": "%d stereo frames (96 images, 48 matches, 10 LocalBA windows) on %d host threads, %.1f s: %s"
                      % (S, threads, t, what)}


def main():
    os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=160, help="stereo frames per step per GPU (160 -> 149 MB of images)")
    ap.add_argument("--ref-frames", type=int, default=24, help="stereo frames per step of the CPU reference arm")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end loop (0: the same K as --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="boundary", choices=["boundary", "all"],
                    help="NCCL all-gather of the shard-boundary left-image record only, or of every left-image record")
    ap.add_argument("--ba-depth", type=int, default=int(os.environ.get("B2S_BA_DEPTH", "1")),
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
