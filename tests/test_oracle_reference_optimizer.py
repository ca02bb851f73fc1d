"""Pins LocalBundleAdjustment and PoseOptimization against the REFERENCE'S OWN SOURCE.

/root/reference/src/Optimizer.cc and src/Converter.cc are compiled in place, unmodified, together with the whole vendored
g2o library (every source Thirdparty/g2o/CMakeLists.txt lists: sparse_optimizer, block_solver, the Levenberg algorithm,
robust kernels, types_six_dof_expmap, linear_solver_eigen / linear_solver_dense ...) -> oracle/_ref/libref_optimizer.so
(oracle/Makefile).  What is NOT the reference's: Eigen (absent from this image; oracle/refshim/eigen/refshim_eigen.h is
an eagerly evaluated stand-in for the API surface g2o uses, SimplicialLDLT = dense LDL^T), the cv stand-in, and the
data-holder Map / KeyFrame / MapPoint / Frame objects (refshim/slam_stubs_optimizer.h) that
oracle/ref_optimizer_glue.cpp fills from the same flattened problem the oracle and the CUDA path take.

So the window selection and graph construction (src/Optimizer.cc:629-856), the 5 + 10 iteration schedule with the outlier
re-classification in between (:858-958), g2o's Levenberg policy (core/optimization_algorithm_levenberg.cpp:61-189:
lambda_0, rho, the 1/3..2/3 / x nu updates, 10 retries, the (iniChi - chi) * 1e3 < iniChi stop rule), buildSystem /
Schur / back-substitution (core/block_solver.hpp), Huber (core/robust_kernel_impl.cpp), the SE3 exponential update
(types/se3quat.h) and the write-back through Converter are the reference's own code.

Every case runs twice (conftest.py `checker`): oracle vs reference source on the CPU, and — marked gpu — the CUDA kernels
vs the reference source directly.  Bar (BASELINE.json north_star): identical accept / reject sequence of the LM trials,
identical outlier sets, pose / landmark deltas within 1e-5 relative (+ 1 float32 ulp of the written value: all sides cast
FP64 -> float like Converter::toCvMat)."""
import ctypes
import os

import numpy as np
import pytest

from oracle_binding import call_local_ba, call_pose_optimization
from synth import synth_local_ba, synth_pose_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_optimizer.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")
RTOL = 1e-5


@pytest.fixture(scope="module")
def ref():
    return ctypes.CDLL(LIB)


def _close(got, want, init):
    got, want, init = (np.asarray(a, np.float64) for a in (got, want, init))
    scale = max(np.abs(want - init).max(), 1e-12)
    ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
    err = np.abs(got - want)
    return bool((err <= RTOL * scale + ulp).all()), float((err / scale).max())


def _check_ba(out, want, d):
    assert out is not None and want is not None
    tr_o = out["trace"][:out["n_trials"]].tolist()
    tr_w = want["trace"][:want["n_trials"]].tolist()
    assert tr_o == tr_w, (tr_o, tr_w)
    assert np.array_equal(out["outlier"], want["outlier"])
    ok, worst = _close(out["Tcw"], want["Tcw"], d["Tcw"][:d["n_local"]])
    assert ok, "pose deltas differ from the reference: %g" % worst
    ok, worst = _close(out["points"], want["points"], d["points"])
    assert ok, "landmark deltas differ from the reference: %g" % worst


def _hard_problem(sd):
    rng = np.random.RandomState(sd)
    ps, ts = rng.uniform(5, 14), rng.uniform(1, 4)
    d = synth_local_ba(n_kf=10, n_fixed=2, n_mp=300, obs_per_mp=5, seed=3, outlier_frac=0.15)
    d["points"] = (d["points"] + rng.normal(0, ps, d["points"].shape)).astype(np.float32)
    T = d["Tcw"].reshape(-1, 4, 4).copy()
    T[1:d["n_local"], :3, 3] += rng.normal(0, ts, (d["n_local"] - 1, 3)).astype(np.float32)
    d["Tcw"] = T.reshape(-1, 16)
    return d


def _shuffled():
    d = synth_local_ba(n_kf=30, n_fixed=6, n_mp=1500, obs_per_mp=5, seed=21)
    rng = np.random.RandomState(5)
    nl = d["n_local"]
    perm = np.arange(d["n_kf"])
    perm[1:nl] = 1 + rng.permutation(nl - 1)
    inv = np.argsort(perm)
    d["Tcw"] = np.ascontiguousarray(d["Tcw"][perm])
    d["fixed"] = np.ascontiguousarray(d["fixed"][perm])
    d["edges"]["kf"] = inv[d["edges"]["kf"]].astype(np.int32)
    return d


BA_CASES = {
    "small": lambda: synth_local_ba(n_kf=8, n_fixed=2, n_mp=200, obs_per_mp=4, seed=5),
    "mixed_mono_stereo": lambda: synth_local_ba(n_kf=12, n_fixed=3, n_mp=600, obs_per_mp=5, seed=7, mono_frac=0.4),
    "all_mono": lambda: synth_local_ba(n_kf=10, n_fixed=3, n_mp=400, obs_per_mp=5, seed=9, mono_frac=1.0),
    "kitti_50_5000_30k": lambda: synth_local_ba(n_kf=50, n_fixed=10, n_mp=5000, obs_per_mp=6, seed=42),
    "rejections_22": lambda: _hard_problem(22),
    "rejections_7": lambda: _hard_problem(7),
    "rejections_2": lambda: _hard_problem(2),
    "shuffled_keyframes": _shuffled,
    "high_degree": lambda: synth_local_ba(n_kf=24, n_fixed=2, n_mp=150, obs_per_mp=20, seed=31),
    "no_outliers": lambda: synth_local_ba(n_kf=10, n_fixed=2, n_mp=300, obs_per_mp=5, seed=15, outlier_frac=0.0),
}


@pytest.mark.parametrize("case", sorted(BA_CASES))
def test_local_ba_matches_reference_source(checker, ref, case):
    d = BA_CASES[case]()
    want = call_local_ba(ref.ref_local_ba, d)
    if case.startswith("rejections_") and case != "rejections_2":
        tr = "".join(map(str, want["trace"][:want["n_trials"]].tolist()))
        assert "01" in tr, tr  # the reference's own LM really rejects and recovers here
    if case == "kitti_50_5000_30k":
        assert want["n_trials"] >= 10 and want["outlier"].sum() > 500
    _check_ba(checker.local_ba(d), want, d)


def test_local_ba_stop_flag_before_round_one(checker, ref):
    """pbStopFlag already set: the reference returns before optimising and writes nothing back (src/Optimizer.cc:858-860)."""
    d = synth_local_ba(n_kf=8, n_fixed=2, n_mp=200, obs_per_mp=4, seed=5)
    stop = np.ones(1, np.uint8)
    assert call_local_ba(ref.ref_local_ba, d, stop) is None
    assert checker.local_ba(d, stop=stop) is None
    stop[0] = 0
    _check_ba(checker.local_ba(d, stop=stop), call_local_ba(ref.ref_local_ba, d, stop), d)


def test_reference_lm_iteration_records(ref):
    """The per-iteration records of the traced Levenberg subclass are consistent with g2o's own rules: lambda shrinks by
    a factor in [1/3, 2/3] after an accepted first trial and the iteration count of a round never exceeds 5 / 10."""
    d = synth_local_ba(n_kf=8, n_fixed=2, n_mp=200, obs_per_mp=4, seed=5)
    call_local_ba(ref.ref_local_ba, d)
    trials = np.zeros(32, np.int32)
    acc = np.zeros(32, np.int32)
    res = np.zeros(32, np.int32)
    lam = np.zeros(32, np.float64)
    ref.ref_lm_iterations.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
    n = ref.ref_lm_iterations(trials.ctypes.data, acc.ctypes.data, res.ctypes.data, lam.ctypes.data, 32)
    assert 2 <= n <= 15 and (trials[:n] >= 1).all() and (lam[:n] > 0).all()
    for i in range(1, min(n, 5)):
        if trials[i] == 1 and acc[i] == 1:
            assert lam[i - 1] / 3 * (1 - 1e-12) <= lam[i] <= lam[i - 1] * 2 / 3 * (1 + 1e-12)


POSE_CASES = [(23, {}), (24, dict(pert_t=0.5, pert_deg=3.0)), (25, dict(outlier_frac=0.4)), (27, dict(mono_frac=1.0)),
              (28, dict(mono_frac=0.0, n=800)), (26, dict(n=12, mp_frac=0.7)), (30, dict(n=300, outlier_frac=0.25)),
              (31, dict(pert_t=1.5, pert_deg=8.0, outlier_frac=0.3))]


@pytest.mark.parametrize("seed,kw", POSE_CASES)
def test_pose_optimization_matches_reference_source(checker, ref, seed, kw):
    d = synth_pose_problem(seed=seed, **kw)
    want = call_pose_optimization(ref.ref_pose_optimization, d)
    out = checker.pose_optimization(d)
    assert out["n_inliers"] == want["n_inliers"]
    assert np.array_equal(out["outlier"], want["outlier"])
    ok, worst = _close(out["Tcw"], want["Tcw"], d["Tcw"])
    assert ok, "pose delta differs from the reference: %g" % worst
    # at convergence the gain of a step is at rounding level, so the last trials may differ between summation orders;
    # the early, well-conditioned trials must agree exactly
    k = min(out["n_trials"], want["n_trials"], 8)
    assert out["trace"][:k].tolist() == want["trace"][:k].tolist()


def test_pose_optimization_no_correspondences(checker, ref):
    d = synth_pose_problem(seed=29, n=40, mp_frac=0.0)  # < 3 correspondences: returns 0 (src/Optimizer.cc:492-493)
    want = call_pose_optimization(ref.ref_pose_optimization, d)
    out = checker.pose_optimization(d)
    assert want["n_inliers"] == 0 and out["n_inliers"] == 0
    assert np.array_equal(out["Tcw"], d["Tcw"]) and not out["outlier"].any()


def _result_equal_ba(out, want, d):
    ok1, _ = _close(out["Tcw"], want["Tcw"], d["Tcw"][:d["n_local"]])
    ok2, _ = _close(out["points"], want["points"], d["points"])
    return ok1 and ok2 and np.array_equal(out["outlier"], want["outlier"])


def test_local_ba_random_sweep(checker, ref):
    """24 random windows (4-20 keyframes, 50-600 points, mono / stereo mixes, half of them strongly perturbed so that LM
    trials are rejected and rounds end early): poses, landmarks and outlier sets must equal the reference's.  The accept /
    reject sequence must be identical up to the point where the robustified chi2 has converged (there the gain of a step
    is at rounding level and its sign depends on the summation order), i.e. on the common prefix minus the last trials."""
    exact = 0
    for sd in range(24):
        rng = np.random.RandomState(1000 + sd)
        nkf = int(rng.randint(4, 20))
        nfix = int(rng.randint(1, max(2, nkf // 3)))
        d = synth_local_ba(n_kf=nkf, n_fixed=nfix, n_mp=int(rng.randint(50, 600)), obs_per_mp=int(rng.randint(2, min(nkf, 8))),
                           seed=sd, mono_frac=float(rng.choice([0, 0.3, 1.0])), outlier_frac=float(rng.choice([0, 0.05, 0.2])))
        if rng.uniform() < 0.5:
            ps, ts = rng.uniform(0.5, 14), rng.uniform(0.1, 4)
            d["points"] = (d["points"] + rng.normal(0, ps, d["points"].shape)).astype(np.float32)
            T = d["Tcw"].reshape(-1, 4, 4).copy()
            T[1:d["n_local"], :3, 3] += rng.normal(0, ts, (d["n_local"] - 1, 3)).astype(np.float32)
            d["Tcw"] = T.reshape(-1, 16)
        want = call_local_ba(ref.ref_local_ba, d)
        out = checker.local_ba(d)
        assert _result_equal_ba(out, want, d), "window %d differs from the reference" % sd
        to, tw = out["trace"][:out["n_trials"]].tolist(), want["trace"][:want["n_trials"]].tolist()
        k = max(min(len(to), len(tw)) - 8, min(len(to), len(tw), 5))
        assert to[:k] == tw[:k], (sd, to, tw)
        exact += to == tw
    assert exact >= 18  # the converged tail differs only now and then


def test_pose_optimization_random_sweep(checker, ref):
    for sd in range(24):
        rng = np.random.RandomState(2000 + sd)
        d = synth_pose_problem(n=int(rng.randint(10, 2500)), seed=sd + 100, mp_frac=float(rng.uniform(0.2, 1)),
                               mono_frac=float(rng.choice([0, 0.2, 1])), outlier_frac=float(rng.uniform(0, 0.5)),
                               pert_t=float(rng.uniform(0.01, 2)), pert_deg=float(rng.uniform(0.1, 10)))
        want = call_pose_optimization(ref.ref_pose_optimization, d)
        out = checker.pose_optimization(d)
        assert out["n_inliers"] == want["n_inliers"] and np.array_equal(out["outlier"], want["outlier"]), sd
        ok, worst = _close(out["Tcw"], want["Tcw"], d["Tcw"])
        assert ok, (sd, worst)
