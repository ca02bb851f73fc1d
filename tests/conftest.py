import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    import importlib
    return importlib.import_module("self_commit_orb-slam2_b200")


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding
    return oracle_binding.load()


class CudaChecker:
    """The CUDA path behind the call signatures of oracle_binding.Oracle, so that the reference-source pinning tests
    (tests/test_oracle_reference_*.py) run twice: oracle vs reference source on the CPU, and — marked gpu — the CUDA
    kernels vs the reference source directly (oracle/_ref/*.so travel to the GPU box with the snapshot)."""
    name = "cuda"

    def __init__(self, pkg):
        self.pkg = pkg

    def search_by_bow(self, descA, nodeA, validA, angA, descB, nodeB, angB, validB=None, th_low=50, nnratio=0.7,
                      strict_lt=False, check_ori=True):
        m = self.pkg.ORBmatcher(nnratio, check_ori)
        return m.SearchByBoW(descA, nodeA, validA, angA, descB, nodeB, angB, validB=validB, strict_lt=strict_lt, th_low=th_low)

    def search_by_projection_last(self, queries, kpx, kpy, octave, angle, uright, occupied, desc, geom, th, mode=0, th_high=100,
                                  check_ori=True):
        m = self.pkg.ORBmatcher(0.9, check_ori)
        return m.SearchByProjection(queries, kpx, kpy, octave, angle, uright, occupied, desc, geom, th, mode=mode, th_high=th_high)

    def search_by_projection_map(self, queries, kpx, kpy, octave, uright, occupied, desc, geom, th=1.0, th_high=100, nnratio=0.8):
        m = self.pkg.ORBmatcher(nnratio, True)
        return m.SearchByProjectionMap(queries, kpx, kpy, octave, uright, occupied, desc, geom, th=th, th_high=th_high)

    def search_for_triangulation(self, kf1, kf2, F12, ex, ey, scale_factors, level_sigma2, only_stereo=False, check_ori=True):
        m = self.pkg.ORBmatcher(0.6, check_ori)
        return m.SearchForTriangulation(kf1, kf2, F12, ex, ey, scale_factors, level_sigma2, only_stereo=only_stereo)

    def search_windows(self, queries, kpx, kpy, octave, uright, inv_level_sigma2, occupied, desc, geom, chi2=False, greedy=False,
                       th_dist=50):
        m = self.pkg.ORBmatcher(0.8, True)
        return m.SearchWindows(queries, kpx, kpy, octave, uright, inv_level_sigma2, occupied, desc, geom, chi2=chi2,
                               greedy=greedy, th_dist=th_dist)

    def distinctive_descriptors(self, desc, offsets):
        return self.pkg.ORBmatcher(0.6, True).DistinctiveDescriptors(desc, offsets)

    def search_for_initialization(self, prev, octave1, angle1, desc1, kpx2, kpy2, octave2, angle2, desc2, geom, window=10,
                                  th_low=50, nnratio=0.9, check_ori=True):
        m = self.pkg.ORBmatcher(nnratio, check_ori)
        return m.SearchForInitialization(prev, octave1, angle1, desc1, kpx2, kpy2, octave2, angle2, desc2, geom, window=window,
                                         th_low=th_low)

    def extractor(self, nfeatures, scaleFactor, nlevels, iniTh, minTh):
        pkg = self.pkg

        def run(img):
            h, w = img.shape
            ex = pkg.ORBextractor(nfeatures, scaleFactor, nlevels, iniTh, minTh, max_width=w, max_height=h, max_batch=1)
            return ex.extract_batch([img])[0]
        return run

    def local_ba(self, d, stop=None):
        opt = self.pkg.Optimizer(max_kf=max(16, d["n_kf"]), max_mp=max(512, len(d["points"])), max_edges=max(4096, len(d["edges"])))
        return opt.LocalBundleAdjustment(d, stop=stop)

    def pose_optimization(self, d):
        return self.pkg.Optimizer(max_kf=4, max_mp=16, max_edges=64).PoseOptimization(d)

    def bow_transform(self, voc, features, levelsup=4):
        v = self.pkg.ORBVocabulary(voc["k"], voc["L"], voc["parent"], voc["leaf_flag"], voc["desc"], voc["weight"])
        word, w, node = v.transform_features(features, levelsup)
        return int((w > 0).sum()), word, w, node


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def checker(request):
    """`oracle` (CPU; the default suite) or the CUDA path (-m gpu) behind the same call signatures."""
    if request.param == "oracle":
        return request.getfixturevalue("oracle")
    return CudaChecker(request.getfixturevalue("pkg"))
