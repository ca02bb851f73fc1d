"""One 32-window LocalBA batch (for ncu captures; B2S_BA_REPEAT=phase*256+count repeats one phase)."""
import sys, time, importlib, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
os.environ.setdefault('B2S_DEBUG_TIMING', '1')
from synth import synth_local_ba
pkg = importlib.import_module('self_commit_orb-slam2_b200')
d = synth_local_ba()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 32
opt = pkg.Optimizer(max_kf=64, max_mp=5000, max_edges=30000, max_batch=nb)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    t = time.perf_counter(); out = opt.LocalBundleAdjustmentBatch([d] * nb); dt = time.perf_counter() - t
    print('batch', nb, 'ms', dt * 1e3, 'trials', out[0]['n_trials'], flush=True)
