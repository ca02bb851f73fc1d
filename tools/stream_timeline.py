"""Where does a pipelined stream step spend its wall time?  Wraps the solver calls of StereoStream with timers and prints, per
step: main-thread enqueue time, time blocked on the LocalBA / PoseOptimization futures, and the wall time of every solver batch.
usage: python tools/stream_timeline.py [steps]"""
import importlib, importlib.util, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
stream_mod = importlib.import_module("self_commit_orb-slam2_b200.stream")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
F, D = 160, 512
windows, poses = b.ba_windows(32), b.pose_problems(F)
ss = stream_mod.StereoStream(F, b.W_IMG, b.H_IMG, b.NFEAT, ba_problems=windows, pose_problems=poses, stereo=True, ba_every=b.BA_EVERY,
                             ba_depth=int(os.environ.get("DEPTH", "2")), bf=b.BF, project=True, intrinsics=(b.FX, b.FY, b.CX, b.CY), motion=b.STREAM_MOTION)
d_all = torch.from_numpy(b.make_stream_images(D)).cuda()
log = []
t00 = time.perf_counter()
def wrap(name, fn):
    def g(*a, **k):
        t = time.perf_counter(); r = fn(*a, **k); log.append((name, t - t00, time.perf_counter() - t00)); return r
    return g
ss._solve = wrap("ba", ss._solve); ss._solve_pose = wrap("pose", ss._solve_pose)
for it in range(steps + 4):
    if it == 4:
        torch.cuda.synchronize(); log.clear(); tstart = time.perf_counter()
    t0 = time.perf_counter()
    ss.load_window(d_all, (it * F) % D)
    with torch.cuda.stream(ss.stream):
        ss._enqueue_extract_match()
    t1 = time.perf_counter()
    ss._submit_ba()
    ss._step_no += 1
    t2 = time.perf_counter()
    if it >= 4:
        log.append(("main", t0 - t00, t1 - t00, t2 - t00))
ss.finish(); torch.cuda.synchronize()
tot = time.perf_counter() - tstart
print("ms/step %.2f" % (tot * 1e3 / steps))
for e in sorted(log, key=lambda e: e[1]):
    if e[0] == "main":
        print("main  start %7.2f enqueue %5.2f ms, blocked in submit %5.2f ms" % ((e[1] - (tstart - t00)) * 1e3, (e[2] - e[1]) * 1e3, (e[3] - e[2]) * 1e3))
    else:
        print("  %-5s start %7.2f dur %6.2f ms" % (e[0], (e[1] - (tstart - t00)) * 1e3, (e[2] - e[1]) * 1e3))
