"""self_commit_orb-slam2_b200 — B200-native hot path of ORB-SLAM2 (extract -> match -> LocalBA).

This package is a thin ctypes binding over ``libb200slam.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/b200slam.h``).  The Python classes mirror the reference's operator interface
(``ORBextractor.__call__`` = ``ORBextractor::operator()``, ``ORBmatcher.SearchByBoW`` ...), so the tests read like
calls into the reference.  There is no CPU fallback: if the CUDA library is missing or no GPU is present the calls
raise ``B200SlamError``.

The directory name contains a dash, so import it with
``importlib.import_module("self_commit_orb-slam2_b200")`` (or ``from b200slam_loader import pkg`` in tests).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200slam.so")

OK, ERR_NO_DEVICE, ERR_BAD_ARG, ERR_CUDA, ERR_CAPACITY, ERR_ABORTED = range(6)

keypoint_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])
win_query_dtype = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("radius", "<f4"), ("min_level", "<i4"),
                            ("max_level", "<i4"), ("valid", "u1"), ("pad", "u1", 3), ("desc", "u1", 32)])
map_query_dtype = np.dtype([("u", "<f4"), ("v", "<f4"), ("ur", "<f4"), ("view_cos", "<f4"), ("level", "<i4"),
                            ("in_view", "u1"), ("has_obs", "u1"), ("pad", "u1", 2), ("desc", "u1", 32)])
proj_query_dtype = np.dtype([("u", "<f4"), ("v", "<f4"), ("invz", "<f4"), ("angle", "<f4"), ("octave", "<i4"),
                             ("has_obs", "<i4"), ("desc", "u1", 32)])
ba_edge_dtype = np.dtype([("kf", "<i4"), ("mp", "<i4"), ("obs", "<f4", 3), ("inv_sigma2", "<f4")])

TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30  # src/ORBmatcher.cc:49-51


class B200SlamError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libb200slam error %d: %s" % (code, text))
        self.code = code


_vp = ctypes.c_void_p
_lib = None


class _FrameGeom(ctypes.Structure):
    _fields_ = [("mnMinX", ctypes.c_float), ("mnMinY", ctypes.c_float), ("mnMaxX", ctypes.c_float),
                ("mnMaxY", ctypes.c_float), ("bf", ctypes.c_float), ("scale_factors", _vp), ("nlevels", ctypes.c_int)]


class _KfFeatures(ctypes.Structure):
    _fields_ = [("desc", ctypes.c_void_p), ("node", ctypes.c_void_p), ("has_mp", ctypes.c_void_p),
                ("stereo", ctypes.c_void_p), ("x", ctypes.c_void_p), ("y", ctypes.c_void_p), ("octave", ctypes.c_void_p),
                ("angle", ctypes.c_void_p), ("n", ctypes.c_int32)]


class _BaProblem(ctypes.Structure):
    _fields_ = [("n_kf", ctypes.c_int), ("n_local", ctypes.c_int), ("Tcw", _vp), ("fixed", _vp), ("n_mp", ctypes.c_int),
                ("points", _vp), ("n_edges", ctypes.c_int), ("edges", _vp), ("fx", ctypes.c_float),
                ("fy", ctypes.c_float), ("cx", ctypes.c_float), ("cy", ctypes.c_float), ("bf", ctypes.c_float),
                ("its1", ctypes.c_int), ("its2", ctypes.c_int)]


class _BaResult(ctypes.Structure):
    _fields_ = [("Tcw_out", _vp), ("points_out", _vp), ("edge_outlier", _vp), ("trace", _vp),
                ("chi2_final", ctypes.c_double), ("n_trials", ctypes.c_int)]


def lib():
    """Loads libb200slam.so (built in-tree by build.py). Fails loudly if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200SlamError(-1, "%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                    "(there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.b2s_last_error.restype = ctypes.c_char_p
        L.b2s_extractor_launch_count.restype = ctypes.c_longlong
        L.b2s_extractor_launch_count.argtypes = [_vp]
        L.b2s_extractor_create.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]
        L.b2s_extractor_destroy.argtypes = [_vp]
        L.b2s_extractor_destroy.restype = None
        L.b2s_extractor_tables.argtypes = [_vp] * 6
        L.b2s_extractor_max_keypoints.argtypes = [_vp]
        L.b2s_extract.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp]
        L.b2s_extract_batch.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp,
                                        ctypes.c_int, _vp]
        L.b2s_extract_batch_device.argtypes = [_vp, _vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, _vp, _vp, _vp, ctypes.c_int, _vp]
        L.b2s_extractor_check.argtypes = [_vp]
        L.b2s_extractor_set_timing.argtypes = [_vp, ctypes.c_int]
        L.b2s_extractor_get_timing.argtypes = [_vp, _vp, _vp]
        L.b2s_extractor_debug_level.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]
        L.b2s_extractor_debug_candidates.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp]
        if hasattr(L, "b2s_matcher_create"):
            L.b2s_matcher_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]
            L.b2s_matcher_destroy.argtypes = [_vp]
            L.b2s_matcher_destroy.restype = None
            L.b2s_matcher_launch_count.restype = ctypes.c_longlong
            L.b2s_matcher_launch_count.argtypes = [_vp]
            L.b2s_descriptor_distance.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp]
            L.b2s_search_by_bow.argtypes = [_vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, _vp, _vp]
            L.b2s_search_by_bow_batch.argtypes = [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp,
                                                  _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                                  ctypes.c_int, _vp, _vp]
            L.b2s_search_by_bow_device.argtypes = [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp,
                                                   _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                   ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]
            L.b2s_search_by_projection_last.argtypes = [_vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                                        ctypes.c_int, _vp, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                        ctypes.c_int, _vp, _vp]
            L.b2s_search_by_projection_last_device.argtypes = [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp,
                                                               ctypes.c_int, _vp, ctypes.c_float, ctypes.c_int,
                                                               ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]
            L.b2s_search_by_projection_last_batch.argtypes = [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp,
                                                              ctypes.c_int, _vp, ctypes.c_float, ctypes.c_int,
                                                              ctypes.c_int, ctypes.c_int, _vp, _vp]
            L.b2s_search_by_projection_sequence.argtypes = [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp,
                                                            ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                                            ctypes.c_int, _vp, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                            ctypes.c_int, _vp, _vp]
            L.b2s_track_queries_device.argtypes = [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_float,
                                                   ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int, _vp, _vp,
                                                   _vp]
            L.b2s_search_by_projection_map.argtypes = [_vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int,
                                                       _vp, ctypes.c_float, ctypes.c_int, ctypes.c_float, _vp, _vp]
        if hasattr(L, "b2s_ba_create"):
            L.b2s_ba_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]
            L.b2s_ba_destroy.argtypes = [_vp]
            L.b2s_ba_destroy.restype = None
            L.b2s_ba_launch_count.restype = ctypes.c_longlong
            L.b2s_ba_launch_count.argtypes = [_vp]
            L.b2s_ba_last_kernel_ms.restype = ctypes.c_float
            L.b2s_ba_last_kernel_ms.argtypes = [_vp, _vp]
            L.b2s_local_ba.argtypes = [_vp, _vp, _vp, _vp]
            L.b2s_ba_set_sm_budget.argtypes = [_vp, ctypes.c_int]
            L.b2s_local_ba_batch.argtypes = [_vp, ctypes.c_int, _vp, _vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != OK:
        raise B200SlamError(rc, lib().b2s_last_error().decode("utf-8", "replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def device_count():
    return lib().b2s_device_count()


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:92-161)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width=1280, max_height=1024,
                 max_batch=1, device=0):
        self._h = _vp()
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.max_batch = max_batch
        _check(lib().b2s_extractor_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height,
                                          max_batch, device, ctypes.byref(self._h)))
        self.cap = lib().b2s_extractor_max_keypoints(self._h)
        s = [np.zeros(nlevels, np.float32) for _ in range(4)]
        nf = np.zeros(nlevels, np.int32)
        _check(lib().b2s_extractor_tables(self._h, _p(s[0]), _p(s[1]), _p(s[2]), _p(s[3]), _p(nf)))
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2 = s
        self.mnFeaturesPerLevel = nf
        self.mvImagePyramid = []

    def close(self):
        if self._h:
            lib().b2s_extractor_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # getters of include/ORBextractor.h:118-161
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactors(self):
        return self.mvScaleFactor

    def GetInverseScaleFactors(self):
        return self.mvInvScaleFactor

    def GetScaleSigmaSquares(self):
        return self.mvLevelSigma2

    def GetInverseScaleSigmaSquares(self):
        return self.mvInvLevelSigma2

    def __call__(self, image, mask=None, want_pyramid=False):
        """operator()(image, mask, keypoints, descriptors): returns (keypoints[N], descriptors[N,32])."""
        if image is None or image.size == 0:
            return np.zeros(0, keypoint_dtype), np.zeros((0, 32), np.uint8)
        assert image.dtype == np.uint8 and image.ndim == 2  # assert(image.type() == CV_8UC1)
        h, w = image.shape
        kps = np.zeros(self.cap, keypoint_dtype)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = ctypes.c_int(0)
        pyr_ptrs = None
        if want_pyramid:
            dims = self.level_dims(w, h)
            self.mvImagePyramid = [np.zeros((lh, lw), np.uint8) for (lw, lh) in dims]
            arr = (_vp * self.nlevels)(*[p.ctypes.data for p in self.mvImagePyramid])
            pyr_ptrs = ctypes.cast(arr, _vp)
        _check(lib().b2s_extract(self._h, image.ctypes.data_as(_vp), w, h, image.strides[0], _p(kps), _p(desc), self.cap,
                                 ctypes.byref(n), pyr_ptrs))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def level_dims(self, w, h):
        return [(int(np.rint(np.float32(w) * s)), int(np.rint(np.float32(h) * s))) for s in self.mvInvScaleFactor]

    def extract_batch(self, images):
        B = len(images)
        h, w = images[0].shape
        kps = np.zeros((B, self.cap), keypoint_dtype)
        desc = np.zeros((B, self.cap, 32), np.uint8)
        n = np.zeros(B, np.int32)
        imgs = [np.ascontiguousarray(im) for im in images]
        ptrs = (_vp * B)(*[im.ctypes.data for im in imgs])
        _check(lib().b2s_extract_batch(self._h, ctypes.cast(ptrs, _vp), B, w, h, w, _p(kps), _p(desc), self.cap, _p(n)))
        return [(kps[b, :n[b]].copy(), desc[b, :n[b]].copy()) for b in range(B)]

    def extract_batch_device(self, d_imgs_ptr, img_pitch_bytes, batch, w, h, stride, d_kps_ptr, d_desc_ptr, d_counts_ptr,
                             cap, stream=None):
        _check(lib().b2s_extract_batch_device(self._h, _vp(d_imgs_ptr), img_pitch_bytes, batch, w, h, stride,
                                              _vp(d_kps_ptr), _vp(d_desc_ptr), _vp(d_counts_ptr), cap,
                                              _vp(stream) if stream else None))

    def stereo_match(self, first_left, first_right, n_pairs, bf, mb=0.0):
        """Frame::ComputeStereoMatches (src/Frame.cc:1026-1421) on the pairs of the batch extracted last with
        extract_batch(); returns (uright [n_pairs, cap], depth [n_pairs, cap], n_matched [n_pairs])."""
        ur = np.full((n_pairs, self.cap), -1, np.float32)
        dp = np.full((n_pairs, self.cap), -1, np.float32)
        nm = np.zeros(n_pairs, np.int32)
        L = lib()
        L.b2s_stereo_match.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, _vp,
                                       _vp, ctypes.c_int, _vp]
        _check(L.b2s_stereo_match(self._h, first_left, first_right, n_pairs, float(bf), float(mb), _p(ur), _p(dp),
                                  self.cap, _p(nm)))
        return ur, dp, nm

    def stereo_match_device(self, first_left, first_right, n_pairs, d_kps_ptr, d_desc_ptr, d_counts_ptr, cap, bf, mb,
                            d_uright_ptr, d_depth_ptr, d_nmatched_ptr, stream=0):
        L = lib()
        L.b2s_stereo_match_device.argtypes = [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, ctypes.c_int,
                                              ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp]
        _check(L.b2s_stereo_match_device(self._h, first_left, first_right, n_pairs, _vp(d_kps_ptr), _vp(d_desc_ptr),
                                         _vp(d_counts_ptr), cap, float(bf), float(mb), _vp(d_uright_ptr),
                                         _vp(d_depth_ptr), _vp(d_nmatched_ptr), _vp(stream)))

    def check(self):
        _check(lib().b2s_extractor_check(self._h))

    def launch_count(self):
        return lib().b2s_extractor_launch_count(self._h)

    # test hooks
    def debug_level(self, b, level, blurred=False):
        w, h = ctypes.c_int(0), ctypes.c_int(0)
        _check(lib().b2s_extractor_debug_level(self._h, b, level, int(blurred), None, ctypes.byref(w), ctypes.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        _check(lib().b2s_extractor_debug_level(self._h, b, level, int(blurred), _p(out), None, None))
        return out

    def debug_candidates(self, b, level):
        cap = 70000
        xy = np.zeros((cap, 2), np.int32)
        resp = np.zeros(cap, np.int32)
        n = ctypes.c_int(0)
        _check(lib().b2s_extractor_debug_candidates(self._h, b, level, _p(xy), _p(resp), cap, ctypes.byref(n)))
        return xy[:n.value].copy(), resp[:n.value].copy()


class ORBmatcher:
    """Mirror of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:57-215) on flattened arrays."""
    TH_HIGH, TH_LOW, HISTO_LENGTH = TH_HIGH, TH_LOW, HISTO_LENGTH

    def __init__(self, nnratio=0.6, checkOri=True, max_features=4096, max_batch=1, device=0):
        self.mfNNratio = np.float32(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self._h = _vp()
        _check(lib().b2s_matcher_create(max_features, max_batch, device, ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().b2s_matcher_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self):
        return lib().b2s_matcher_launch_count(self._h)

    def DescriptorDistance(self, a, b):
        a = np.ascontiguousarray(a, np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        _check(lib().b2s_descriptor_distance(self._h, _p(a), _p(b), len(a), _p(out)))
        return out

    def SearchByBoW(self, descA, nodeA, validA, angA, descB, nodeB, angB, validB=None, strict_lt=False, th_low=TH_LOW):
        nA, nB = len(descA), len(descB)
        match = np.full(nB, -1, np.int32)
        nm = ctypes.c_int(0)
        _check(lib().b2s_search_by_bow(self._h, _p(descA), _p(nodeA), _p(validA), _p(angA), nA, _p(descB), _p(nodeB),
                                       _p(validB), _p(angB), nB, th_low, float(self.mfNNratio), int(strict_lt),
                                       int(self.mbCheckOrientation), _p(match), ctypes.byref(nm)))
        return nm.value, match

    def SearchByProjection(self, queries, kpx, kpy, octave, angle, uright, occupied, desc, geom, th, mode=0,
                           th_high=TH_HIGH):
        nf = len(kpx)
        sf = np.ascontiguousarray(geom["scale_factors"], np.float32)
        g = _FrameGeom(geom["mnMinX"], geom["mnMinY"], geom["mnMaxX"], geom["mnMaxY"], geom["bf"], sf.ctypes.data, len(sf))
        match = np.full(nf, -1, np.int32)
        nm = ctypes.c_int(0)
        _check(lib().b2s_search_by_projection_last(self._h, _p(queries), len(queries), _p(kpx), _p(kpy), _p(octave),
                                                   _p(angle), _p(uright), _p(occupied), _p(desc), nf, ctypes.byref(g),
                                                   float(th), mode, th_high, int(self.mbCheckOrientation), _p(match),
                                                   ctypes.byref(nm)))
        return nm.value, match


    @staticmethod
    def frame_geom(geom):
        """ctypes b2s_frame_geom (+ the array it points to, which the caller keeps alive) from a dict."""
        sf = np.ascontiguousarray(geom["scale_factors"], np.float32)
        return _FrameGeom(geom["mnMinX"], geom["mnMinY"], geom["mnMaxX"], geom["mnMaxY"], geom["bf"], sf.ctypes.data,
                          len(sf)), sf

    def track_queries_device(self, batch, d_kps_last, d_desc_last, d_depth_last, d_n_last, cap, d_Tcl, fx, fy, cx, cy, has_obs,
                             d_q, d_nq, stream=None):
        """Projection half of SearchByProjection(CurrentFrame, LastFrame) (src/ORBmatcher.cc:1600-1626) for `batch` stereo
        frame pairs on the device (b2s_track_queries_device); all d_* are device addresses."""
        _check(lib().b2s_track_queries_device(self._h, batch, _vp(d_kps_last), _vp(d_desc_last), _vp(d_depth_last),
                                              _vp(d_n_last), cap, _vp(d_Tcl), float(fx), float(fy), float(cx), float(cy),
                                              int(has_obs), _vp(d_q), _vp(d_nq), _vp(stream) if stream else None))

    def search_by_projection_last_device(self, batch, d_q, d_nq, cap_q, d_kps, d_uright, d_desc, d_nf, cap_f, geom, th, mode,
                                         d_match, d_nmatches, th_high=TH_HIGH, stream=None):
        """Batched device-resident SearchByProjection(CurrentFrame, LastFrame) (b2s_search_by_projection_last_device)."""
        g, keep = self.frame_geom(geom)
        _check(lib().b2s_search_by_projection_last_device(self._h, batch, _vp(d_q), _vp(d_nq), cap_q, _vp(d_kps), _vp(d_uright),
                                                          _vp(d_desc), _vp(d_nf), cap_f, ctypes.byref(g), float(th), int(mode),
                                                          th_high, int(self.mbCheckOrientation), _vp(d_match),
                                                          _vp(d_nmatches), _vp(stream) if stream else None))

    def SearchByProjectionBatch(self, queries, nq, kps, uright, desc, nf, geom, th, mode=0, th_high=TH_HIGH):
        """Host-buffer batch (b2s_search_by_projection_last_batch): queries [B, capQ] proj_query_dtype, kps [B, capF]
        keypoint_dtype, uright [B, capF] f32, desc [B, capF, 32] u8.  Returns (nmatches [B], match_cur [B, capF])."""
        queries, kps = np.ascontiguousarray(queries), np.ascontiguousarray(kps)
        uright, desc = np.ascontiguousarray(uright, np.float32), np.ascontiguousarray(desc, np.uint8)
        B, capQ = queries.shape
        capF = kps.shape[1]
        g, keep = self.frame_geom(geom)
        match = np.full((B, capF), -1, np.int32)
        nm = np.zeros(B, np.int32)
        nq = np.ascontiguousarray(nq, np.int32)
        nf = np.ascontiguousarray(nf, np.int32)
        _check(lib().b2s_search_by_projection_last_batch(self._h, B, _p(queries), _p(nq), capQ, _p(kps), _p(uright), _p(desc),
                                                         _p(nf), capF, ctypes.byref(g), float(th), int(mode), th_high,
                                                         int(self.mbCheckOrientation), _p(match), _p(nm)))
        return nm, match

    def SearchByProjectionSequence(self, kps, desc, depth, uright, n, Tcl, fx, fy, cx, cy, geom, th, mode=0, has_obs=1,
                                   th_high=TH_HIGH, out=None):
        """b2s_search_by_projection_sequence: kps / desc / depth of B + 1 consecutive frames ([B + 1, cap, ...], C-contiguous
        host arrays), uright [B, cap] of frames 1..B, Tcl [B, 12]; frame b + 1 is matched against frame b.  Returns
        (nmatches [B], match_cur [B, cap]) (written into `out` = (nm, match) when given)."""
        B, cap = uright.shape
        g, keep = self.frame_geom(geom)
        nm, match = out if out is not None else (np.zeros(B, np.int32), np.full((B, cap), -1, np.int32))
        _check(lib().b2s_search_by_projection_sequence(self._h, B, _p(kps), _p(desc), _p(depth), _p(uright), _p(n), cap, _p(Tcl),
                                                       float(fx), float(fy), float(cx), float(cy), int(has_obs),
                                                       ctypes.byref(g), float(th), int(mode), th_high,
                                                       int(self.mbCheckOrientation), _p(match), _p(nm)))
        return nm, match

    def SearchForInitialization(self, prev, octave1, angle1, desc1, kpx2, kpy2, octave2, angle2, desc2, geom, window=10,
                                th_low=TH_LOW):
        """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:515-643).  prev: [n1, 2] vbPrevMatched (updated in place
        like the reference does, :636-638).  Returns (nmatches, vnMatches12)."""
        n1, n2 = len(octave1), len(kpx2)
        prev_in = np.ascontiguousarray(prev, np.float32)
        px, py = np.ascontiguousarray(prev_in[:, 0]), np.ascontiguousarray(prev_in[:, 1])
        sf = np.ascontiguousarray(geom["scale_factors"], np.float32)
        g = _FrameGeom(geom["mnMinX"], geom["mnMinY"], geom["mnMaxX"], geom["mnMaxY"], geom["bf"], sf.ctypes.data, len(sf))
        keep = [np.ascontiguousarray(octave1, np.int32), np.ascontiguousarray(angle1, np.float32),
                np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(kpx2, np.float32),
                np.ascontiguousarray(kpy2, np.float32), np.ascontiguousarray(octave2, np.int32),
                np.ascontiguousarray(angle2, np.float32), np.ascontiguousarray(desc2, np.uint8)]
        m12 = np.full(n1, -1, np.int32)
        nm = ctypes.c_int(0)
        L = lib()
        L.b2s_search_for_initialization.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp,
                                                    ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                    ctypes.c_int, _vp, _vp]
        _check(L.b2s_search_for_initialization(self._h, _p(px), _p(py), _p(keep[0]), _p(keep[1]), _p(keep[2]), n1, _p(keep[3]),
                                               _p(keep[4]), _p(keep[5]), _p(keep[6]), _p(keep[7]), n2, ctypes.byref(g),
                                               int(window), int(th_low), float(self.mfNNratio),
                                               int(self.mbCheckOrientation), _p(m12), ctypes.byref(nm)))
        hit = m12 >= 0
        prev[hit, 0] = keep[3][m12[hit]]
        prev[hit, 1] = keep[4][m12[hit]]
        return nm.value, m12

    def SearchByProjectionReloc(self, queries, kpx, kpy, octave, angle, occupied, desc, geom, th, orb_dist):
        """SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1731-1862, relocalisation):
        queries = the keyframe's projected map points with octave = predicted level; occupied[j] = the frame feature holds
        any map point.  Same kernel as the last-frame search: levels [l-1, l+1], no stereo gate, th_high = ORBdist."""
        q = np.ascontiguousarray(queries).copy()
        q["invz"] = 1.0
        q["has_obs"] = 1
        return self.SearchByProjection(q, kpx, kpy, octave, angle, np.full(len(kpx), -1, np.float32), occupied, desc, geom, th,
                                       mode=0, th_high=orb_dist)

    def SearchByProjectionMap(self, queries, kpx, kpy, octave, uright, occupied, desc, geom, th=1.0, th_high=TH_HIGH):
        """SearchByProjection(Frame&, const vector<MapPoint*>&, th) (src/ORBmatcher.cc:70-175); queries: map_query_dtype."""
        nf = len(kpx)
        sf = np.ascontiguousarray(geom["scale_factors"], np.float32)
        g = _FrameGeom(geom["mnMinX"], geom["mnMinY"], geom["mnMaxX"], geom["mnMaxY"], geom["bf"], sf.ctypes.data, len(sf))
        match = np.full(nf, -1, np.int32)
        nm = ctypes.c_int(0)
        _check(lib().b2s_search_by_projection_map(self._h, _p(queries), len(queries), _p(kpx), _p(kpy), _p(octave),
                                                  _p(uright), _p(occupied), _p(desc), nf, ctypes.byref(g), float(th),
                                                  th_high, float(self.mfNNratio), _p(match), ctypes.byref(nm)))
        return nm.value, match


    def SearchWindows(self, queries, kpx, kpy, octave, uright, inv_level_sigma2, occupied, desc, geom, chi2=False,
                      greedy=False, th_dist=TH_LOW):
        """Search core of Fuse (src/ORBmatcher.cc:1020-1174, :1179-1310) and SearchByProjection(KeyFrame*, Scw, ...)
        (:388-512); queries: win_query_dtype.  Returns (n_accepted, best_idx, best_dist)."""
        nq, nf = len(queries), len(kpx)
        sf = np.ascontiguousarray(geom["scale_factors"], np.float32)
        g = _FrameGeom(geom["mnMinX"], geom["mnMinY"], geom["mnMaxX"], geom["mnMaxY"], geom["bf"], sf.ctypes.data, len(sf))
        best = np.full(nq, -1, np.int32)
        bdist = np.full(nq, 256, np.int32)
        nacc = ctypes.c_int(0)
        L = lib()
        L.b2s_search_windows.argtypes = [_vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp,
                                         ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]
        is2 = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
        _check(L.b2s_search_windows(self._h, _p(queries), nq, _p(kpx), _p(kpy), _p(octave), _p(uright), _p(is2),
                                    _p(occupied), _p(desc), nf, ctypes.byref(g), (1 if chi2 else 0) | (2 if greedy else 0),
                                    th_dist, _p(best), _p(bdist), ctypes.byref(nacc)))
        return nacc.value, best, bdist


    def DistinctiveDescriptors(self, desc, offsets):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359-440) for a batch of map points."""
        desc = np.ascontiguousarray(desc, np.uint8)
        offsets = np.ascontiguousarray(offsets, np.int32)
        best = np.full(len(offsets) - 1, -1, np.int32)
        L = lib()
        L.b2s_distinctive_descriptors.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp]
        _check(L.b2s_distinctive_descriptors(self._h, _p(desc), _p(offsets), len(offsets) - 1, _p(best)))
        return best

    def SearchForTriangulation(self, kf1, kf2, F12, ex, ey, scale_factors, level_sigma2, only_stereo=False):
        """ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:810-1009). kf1/kf2: dicts with desc, node, has_mp, stereo,
        x, y, octave, angle.  Returns (nmatches, match12)."""
        keep = []

        def side(k):
            arrs = [np.ascontiguousarray(k["desc"], np.uint8), np.ascontiguousarray(k["node"], np.int32),
                    np.ascontiguousarray(k["has_mp"], np.uint8), np.ascontiguousarray(k["stereo"], np.uint8),
                    np.ascontiguousarray(k["x"], np.float32), np.ascontiguousarray(k["y"], np.float32),
                    np.ascontiguousarray(k["octave"], np.int32), np.ascontiguousarray(k["angle"], np.float32)]
            keep.append(arrs)
            return _KfFeatures(*[a.ctypes.data for a in arrs], len(arrs[0]))

        a, b = side(kf1), side(kf2)
        F = np.ascontiguousarray(F12, np.float32).reshape(9)
        sf = np.ascontiguousarray(scale_factors, np.float32)
        s2 = np.ascontiguousarray(level_sigma2, np.float32)
        match12 = np.full(a.n, -1, np.int32)
        nm = ctypes.c_int(0)
        L = lib()
        L.b2s_search_for_triangulation.argtypes = [_vp, _vp, _vp, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp]
        _check(L.b2s_search_for_triangulation(self._h, ctypes.byref(a), ctypes.byref(b), _p(F), float(ex), float(ey), _p(sf),
                                              _p(s2), len(sf), int(only_stereo), int(self.mbCheckOrientation),
                                              _p(match12), ctypes.byref(nm)))
        return nm.value, match12


class _VocDesc(ctypes.Structure):
    _fields_ = [("k", ctypes.c_int32), ("L", ctypes.c_int32), ("n_nodes", ctypes.c_int32), ("parent", ctypes.c_void_p),
                ("leaf_flag", ctypes.c_void_p), ("desc", ctypes.c_void_p), ("weight", ctypes.c_void_p)]


class FrameGrid:
    """Frame::AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cc:461-491, 741-852) kept on the device: build once per
    frame, then run several projection matchers that upload only their queries."""

    def __init__(self, matcher, kpx=None, kpy=None, octave=None, angle=None, uright=None, desc=None, geom=None,
                 inv_level_sigma2=None, d_kps_ptr=None, d_desc_ptr=None, d_uright_ptr=None, nf=None, stream=0):
        self.matcher = matcher
        self._h = _vp()
        L = lib()
        L.b2s_frame_grid_create.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp]
        L.b2s_frame_grid_create_device.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]
        L.b2s_frame_grid_destroy.argtypes = [_vp]
        L.b2s_frame_grid_destroy.restype = None
        L.b2s_frame_grid_size.argtypes = [_vp]
        L.b2s_frame_grid_features_in_area.argtypes = [_vp, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_int,
                                                      ctypes.c_int, _vp, ctypes.c_int, _vp]
        L.b2s_search_by_projection_last_grid.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_float, ctypes.c_int,
                                                         ctypes.c_int, ctypes.c_int, _vp, _vp]
        L.b2s_search_by_projection_map_grid.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_float, ctypes.c_int,
                                                        ctypes.c_float, _vp, _vp]
        L.b2s_search_windows_grid.argtypes = [_vp, _vp, _vp, ctypes.c_int, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]
        sf = np.ascontiguousarray(geom["scale_factors"], np.float32)
        g = _FrameGeom(geom["mnMinX"], geom["mnMinY"], geom["mnMaxX"], geom["mnMaxY"], geom["bf"], sf.ctypes.data, len(sf))
        is2 = None if inv_level_sigma2 is None else np.ascontiguousarray(inv_level_sigma2, np.float32)
        if d_kps_ptr is not None:
            _check(L.b2s_frame_grid_create_device(matcher._h, _vp(d_kps_ptr), _vp(d_desc_ptr), int(nf),
                                                  _vp(d_uright_ptr) if d_uright_ptr else None, ctypes.byref(g), _p(is2),
                                                  _vp(stream), ctypes.byref(self._h)))
        else:
            a = [np.ascontiguousarray(kpx, np.float32), np.ascontiguousarray(kpy, np.float32),
                 np.ascontiguousarray(octave, np.int32), np.ascontiguousarray(angle, np.float32),
                 np.ascontiguousarray(uright, np.float32), np.ascontiguousarray(desc, np.uint8)]
            _check(L.b2s_frame_grid_create(matcher._h, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]), len(a[0]),
                                           ctypes.byref(g), _p(is2), ctypes.byref(self._h)))
        self.n = lib().b2s_frame_grid_size(self._h)

    def close(self):
        if self._h:
            lib().b2s_frame_grid_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        out = np.zeros(max(self.n, 1), np.int32)
        n = ctypes.c_int(0)
        _check(lib().b2s_frame_grid_features_in_area(self._h, float(x), float(y), float(r), int(minLevel), int(maxLevel),
                                                     _p(out), len(out), ctypes.byref(n)))
        return out[:n.value].copy()

    def SearchByProjection(self, queries, occupied, th, mode=0, th_high=TH_HIGH):
        m = np.full(self.n, -1, np.int32)
        nm = ctypes.c_int(0)
        q = np.ascontiguousarray(queries)
        occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        _check(lib().b2s_search_by_projection_last_grid(self.matcher._h, self._h, _p(q), len(q), _p(occ), float(th), int(mode),
                                                        int(th_high), int(self.matcher.mbCheckOrientation), _p(m),
                                                        ctypes.byref(nm)))
        return nm.value, m

    def SearchByProjectionMap(self, queries, occupied, th=1.0, th_high=TH_HIGH):
        m = np.full(self.n, -1, np.int32)
        nm = ctypes.c_int(0)
        q = np.ascontiguousarray(queries)
        occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        _check(lib().b2s_search_by_projection_map_grid(self.matcher._h, self._h, _p(q), len(q), _p(occ), float(th), int(th_high),
                                                       float(self.matcher.mfNNratio), _p(m), ctypes.byref(nm)))
        return nm.value, m

    def SearchWindows(self, queries, occupied, chi2=False, greedy=False, th_dist=TH_LOW):
        q = np.ascontiguousarray(queries)
        best = np.full(len(q), -1, np.int32)
        bd = np.full(len(q), 256, np.int32)
        na = ctypes.c_int(0)
        occ = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
        _check(lib().b2s_search_windows_grid(self.matcher._h, self._h, _p(q), len(q), _p(occ), (1 if chi2 else 0) | (2 if greedy else 0),
                                             int(th_dist), _p(best), _p(bd), ctypes.byref(na)))
        return na.value, best, bd


class ORBVocabulary:
    """Mirror of ORBVocabulary::transform (DBoW2 TemplatedVocabulary.h:1127-1256) on a flattened vocabulary."""

    def __init__(self, k, L, parent, leaf_flag, desc, weight, device=0):
        self._h = _vp()
        self._keep = [np.ascontiguousarray(parent, np.int32), np.ascontiguousarray(leaf_flag, np.uint8),
                      np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(weight, np.float64)]
        d = _VocDesc(k, L, len(self._keep[0]), *[a.ctypes.data for a in self._keep])
        L_ = lib()
        L_.b2s_vocabulary_create.argtypes = [_vp, ctypes.c_int, _vp]
        L_.b2s_vocabulary_destroy.argtypes = [_vp]
        L_.b2s_vocabulary_destroy.restype = None
        L_.b2s_bow_transform.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]
        L_.b2s_bow_transform_device.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp, _vp]
        _check(L_.b2s_vocabulary_create(ctypes.byref(d), device, ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().b2s_vocabulary_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform_features(self, features, levelsup=4):
        """Per feature (word_id, weight, node_id)."""
        features = np.ascontiguousarray(features, np.uint8)
        n = len(features)
        word = np.zeros(n, np.int32)
        w = np.zeros(n, np.float64)
        node = np.zeros(n, np.int32)
        _check(lib().b2s_bow_transform(self._h, _p(features), n, levelsup, _p(word), _p(w), _p(node)))
        return word, w, node

    def transform(self, features, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup) (:1127-1187) for TF_IDF / L1: returns
        (bow: dict word -> value, feat: dict node -> [feature indices])."""
        word, w, node = self.transform_features(features, levelsup)
        bow, feat = {}, {}
        for i in range(len(word)):
            if w[i] > 0:  # not stopped
                bow[int(word[i])] = bow.get(int(word[i]), 0.0) + float(w[i])  # BowVector::addWeight, feature order
                feat.setdefault(int(node[i]), []).append(i)                   # FeatureVector::addFeature
        norm = 0.0
        for k in sorted(bow):  # BowVector::normalize(L1) iterates the std::map in key order
            norm += abs(bow[k])
        if norm > 0.0:
            for k in bow:
                bow[k] /= norm
        return bow, feat


class Optimizer:
    """Mirror of ORB_SLAM2::Optimizer::LocalBundleAdjustment (include/Optimizer.h:112) on a flattened window."""

    def __init__(self, max_kf=64, max_mp=8192, max_edges=65536, max_batch=1, device=0):
        self._h = _vp()
        _check(lib().b2s_ba_create(max_kf, max_mp, max_edges, max_batch, device, ctypes.byref(self._h)))

    def close(self):
        if self._h:
            lib().b2s_ba_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch_count(self):
        return lib().b2s_ba_launch_count(self._h)

    def set_sm_budget(self, sms):
        """SMs one batched LocalBA call may occupy (b2s_ba_set_sm_budget); 0 = all."""
        _check(lib().b2s_ba_set_sm_budget(self._h, int(sms)))

    def last_kernel_ms(self):
        """(ms, trials): duration of the persistent LM kernel of the last LocalBundleAdjustment(Batch) call, measured with
        CUDA events on the solver's stream, and the LM trials it ran over all windows."""
        n = ctypes.c_longlong(0)
        ms = lib().b2s_ba_last_kernel_ms(self._h, ctypes.byref(n))
        return float(ms), int(n.value)

    @staticmethod
    def _problem(d, its1=5, its2=10):
        keep = dict(Tcw=np.ascontiguousarray(d["Tcw"], np.float32), fixed=np.ascontiguousarray(d["fixed"], np.uint8),
                    points=np.ascontiguousarray(d["points"], np.float32), edges=np.ascontiguousarray(d["edges"]))
        assert keep["edges"].dtype == ba_edge_dtype
        p = _BaProblem(d["n_kf"], d["n_local"], keep["Tcw"].ctypes.data, keep["fixed"].ctypes.data, len(keep["points"]),
                       keep["points"].ctypes.data, len(keep["edges"]), keep["edges"].ctypes.data, d["fx"], d["fy"],
                       d["cx"], d["cy"], d["bf"], its1, its2)
        return p, keep

    @staticmethod
    def _result(d):
        out = dict(Tcw=np.zeros((d["n_local"], 16), np.float32), points=np.zeros((len(d["points"]), 3), np.float32),
                   outlier=np.zeros(len(d["edges"]), np.uint8), trace=np.full(256, -1, np.int32))
        r = _BaResult(out["Tcw"].ctypes.data, out["points"].ctypes.data, out["outlier"].ctypes.data,
                      out["trace"].ctypes.data, 0.0, 0)
        return r, out

    def LocalBundleAdjustment(self, d, stop=None, its1=5, its2=10):
        p, keep = self._problem(d, its1, its2)
        r, out = self._result(d)
        rc = lib().b2s_local_ba(self._h, ctypes.byref(p), _p(stop), ctypes.byref(r))
        if rc == ERR_ABORTED:
            return None
        _check(rc)
        out["chi2"] = r.chi2_final
        out["n_trials"] = r.n_trials
        return out

    def PoseOptimizationBatch(self, ds):
        """Optimizer::PoseOptimization (src/Optimizer.cc:363-605) for a batch of independent frames; ds: dicts with Tcw,
        has_mp, Xw, kpx, kpy, uright, inv_sigma2, fx, fy, cx, cy, bf.  Returns a list of result dicts."""
        class PP(ctypes.Structure):
            _fields_ = [("Tcw", _vp), ("n", ctypes.c_int32), ("has_mp", _vp), ("Xw", _vp), ("kpx", _vp), ("kpy", _vp),
                        ("uright", _vp), ("inv_sigma2", _vp), ("fx", ctypes.c_float), ("fy", ctypes.c_float),
                        ("cx", ctypes.c_float), ("cy", ctypes.c_float), ("bf", ctypes.c_float)]

        class PR(ctypes.Structure):
            _fields_ = [("Tcw_out", _vp), ("outlier", _vp), ("trace", _vp), ("n_inliers", ctypes.c_int32),
                        ("n_trials", ctypes.c_int32)]
        B = len(ds)
        probs, ress, keep, outs = (PP * B)(), (PR * B)(), [], []
        for i, d in enumerate(ds):
            a = dict(Tcw=np.ascontiguousarray(d["Tcw"], np.float32).reshape(16),
                     has_mp=np.ascontiguousarray(d["has_mp"], np.uint8), Xw=np.ascontiguousarray(d["Xw"], np.float32),
                     kpx=np.ascontiguousarray(d["kpx"], np.float32), kpy=np.ascontiguousarray(d["kpy"], np.float32),
                     uright=np.ascontiguousarray(d["uright"], np.float32),
                     inv_sigma2=np.ascontiguousarray(d["inv_sigma2"], np.float32))
            n = len(a["has_mp"])
            o = dict(Tcw=np.zeros(16, np.float32), outlier=np.zeros(max(n, 1), np.uint8), trace=np.full(256, -1, np.int32))
            keep.append((a, o))
            probs[i] = PP(a["Tcw"].ctypes.data, n, a["has_mp"].ctypes.data, a["Xw"].ctypes.data, a["kpx"].ctypes.data,
                          a["kpy"].ctypes.data, a["uright"].ctypes.data, a["inv_sigma2"].ctypes.data, float(d["fx"]),
                          float(d["fy"]), float(d["cx"]), float(d["cy"]), float(d["bf"]))
            ress[i] = PR(o["Tcw"].ctypes.data, o["outlier"].ctypes.data, o["trace"].ctypes.data, 0, 0)
            outs.append((o, n))
        L = lib()
        L.b2s_pose_optimization_batch.argtypes = [_vp, ctypes.c_int, _vp, _vp]
        _check(L.b2s_pose_optimization_batch(self._h, B, ctypes.cast(probs, _vp), ctypes.cast(ress, _vp)))
        return [dict(n_inliers=ress[i].n_inliers, n_trials=ress[i].n_trials, Tcw=o["Tcw"], outlier=o["outlier"][:n],
                     trace=o["trace"]) for i, (o, n) in enumerate(outs)]

    def PoseOptimization(self, d):
        return self.PoseOptimizationBatch([d])[0]

    # ---- prepared batches: the ctypes problem / result arrays are built once, a step is then one C call (stream.py)
    def prepare_pose_batch(self, ds):
        class PP(ctypes.Structure):
            _fields_ = [("Tcw", _vp), ("n", ctypes.c_int32), ("has_mp", _vp), ("Xw", _vp), ("kpx", _vp), ("kpy", _vp),
                        ("uright", _vp), ("inv_sigma2", _vp), ("fx", ctypes.c_float), ("fy", ctypes.c_float),
                        ("cx", ctypes.c_float), ("cy", ctypes.c_float), ("bf", ctypes.c_float)]

        class PR(ctypes.Structure):
            _fields_ = [("Tcw_out", _vp), ("outlier", _vp), ("trace", _vp), ("n_inliers", ctypes.c_int32),
                        ("n_trials", ctypes.c_int32)]
        B = len(ds)
        probs, ress, keep = (PP * B)(), (PR * B)(), []
        for i, d in enumerate(ds):
            a = [np.ascontiguousarray(d["Tcw"], np.float32).reshape(16), np.ascontiguousarray(d["has_mp"], np.uint8),
                 np.ascontiguousarray(d["Xw"], np.float32), np.ascontiguousarray(d["kpx"], np.float32),
                 np.ascontiguousarray(d["kpy"], np.float32), np.ascontiguousarray(d["uright"], np.float32),
                 np.ascontiguousarray(d["inv_sigma2"], np.float32)]
            n = len(a[1])
            o = [np.zeros(16, np.float32), np.zeros(max(n, 1), np.uint8)]
            keep.append((a, o))
            probs[i] = PP(a[0].ctypes.data, n, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, a[4].ctypes.data,
                          a[5].ctypes.data, a[6].ctypes.data, float(d["fx"]), float(d["fy"]), float(d["cx"]), float(d["cy"]),
                          float(d["bf"]))
            ress[i] = PR(o[0].ctypes.data, o[1].ctypes.data, None, 0, 0)
        L = lib()
        L.b2s_pose_optimization_batch.argtypes = [_vp, ctypes.c_int, _vp, _vp]
        return dict(B=B, probs=probs, ress=ress, keep=keep)

    def run_prepared_pose(self, prep):
        _check(lib().b2s_pose_optimization_batch(self._h, prep["B"], ctypes.cast(prep["probs"], _vp),
                                                 ctypes.cast(prep["ress"], _vp)))
        return prep["ress"]

    def prepare_ba_batch(self, ds, its1=5, its2=10):
        B = len(ds)
        probs, ress, keeps, outs = (_BaProblem * B)(), (_BaResult * B)(), [], []
        for i, d in enumerate(ds):
            p, keep = self._problem(d, its1, its2)
            r, out = self._result(d)
            probs[i], ress[i] = p, r
            keeps.append(keep)
            outs.append(out)
        return dict(B=B, probs=probs, ress=ress, keeps=keeps, outs=outs)

    def run_prepared_ba(self, prep):
        _check(lib().b2s_local_ba_batch(self._h, prep["B"], ctypes.cast(prep["probs"], _vp), ctypes.cast(prep["ress"], _vp)))
        for i in range(prep["B"]):
            prep["outs"][i]["chi2"] = prep["ress"][i].chi2_final
            prep["outs"][i]["n_trials"] = prep["ress"][i].n_trials
        return prep["outs"]

    def LocalBundleAdjustmentBatch(self, ds, its1=5, its2=10):
        B = len(ds)
        probs = (_BaProblem * B)()
        ress = (_BaResult * B)()
        keeps, outs = [], []
        for i, d in enumerate(ds):
            p, keep = self._problem(d, its1, its2)
            r, out = self._result(d)
            probs[i], ress[i] = p, r
            keeps.append(keep)
            outs.append(out)
        _check(lib().b2s_local_ba_batch(self._h, B, ctypes.cast(probs, _vp), ctypes.cast(ress, _vp)))
        for i in range(B):
            outs[i]["chi2"] = ress[i].chi2_final
            outs[i]["n_trials"] = ress[i].n_trials
        return outs
