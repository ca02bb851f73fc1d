"""Pins the oracle's OpenCV-owned stages bit-for-bit against cv2 (the same OpenCV the reference links)
and its derived constants against the reference constructor arithmetic (src/ORBextractor.cc:492-609)."""
import numpy as np
import pytest

from synth import synth_image

cv2 = pytest.importorskip("cv2")

SIZES = [(1241, 376), (640, 480), (333, 222), (179, 134)]


@pytest.mark.parametrize("w,h", SIZES)
def test_resize_matches_cv2(oracle, w, h):
    src = synth_image(w, h, 7)
    for s in (1.2, 1.2 ** 2, 1.37):
        dw, dh = int(round(w / s)), int(round(h / s))
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(oracle.resize(src, dw, dh), ref)


@pytest.mark.parametrize("w,h", SIZES)
def test_blur_matches_cv2(oracle, w, h):
    src = synth_image(w, h, 8)
    ref = cv2.GaussianBlur(src, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(oracle.blur(src), ref)


@pytest.mark.parametrize("w,h", SIZES[:3])
@pytest.mark.parametrize("th", [20, 7])
def test_fast_matches_cv2(oracle, w, h, th):
    src = synth_image(w, h, 9)
    det = cv2.FastFeatureDetector_create(th, True)
    ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(src)]
    xy, rs = oracle.fast(src, th)
    mine = [(int(x), int(y), int(r)) for (x, y), r in zip(xy, rs)]
    assert mine == ref and len(ref) > 100


def test_fast_on_cell_sized_rois(oracle):
    """cv::FAST on 37x38 sub-images (the per-cell call of src/ORBextractor.cc:1126)."""
    img = synth_image(640, 480, 10)
    det20 = cv2.FastFeatureDetector_create(20, True)
    for (x0, y0) in [(16, 16), (47, 16), (300, 200), (590, 430)]:
        roi = np.ascontiguousarray(img[y0:y0 + 38, x0:x0 + 37])
        ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det20.detect(roi)]
        xy, rs = oracle.fast(roi, 20)
        assert [(int(x), int(y), int(r)) for (x, y), r in zip(xy, rs)] == ref


def test_fast_atan2_matches_cv2(oracle):
    rng = np.random.RandomState(1)
    ys = rng.randint(-200000, 200000, size=20000).astype(np.float32)
    xs = rng.randint(-200000, 200000, size=20000).astype(np.float32)
    ys[:10] = 0
    xs[5:15] = 0
    for y, x in zip(ys, xs):
        assert np.float32(oracle.fast_atan2(y, x)) == np.float32(cv2.fastAtan2(float(y), float(x)))


def test_constructor_tables(oracle):
    """per-level quotas, level sizes and umax for the two reference configurations."""
    e1 = oracle.extractor(1000, 1.2, 8, 20, 7)
    (_, nf, um) = e1.tables()
    assert nf.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert um.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    e2 = oracle.extractor(2000, 1.2, 8, 20, 7)
    (t, nf2, _) = e2.tables()
    assert nf2.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    img = synth_image(1241, 376, 0)
    e2(img)
    dims = [e2.level(l).shape[::-1] for l in range(8)]
    assert dims == [(1241, 376), (1034, 313), (862, 261), (718, 218), (598, 181), (499, 151), (416, 126), (346, 105)]
    e1(synth_image(640, 480, 0))
    dims = [e1.level(l).shape[::-1] for l in range(8)]
    assert dims == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]


def test_pyramid_chain_and_blur_match_cv2_inside_extractor(oracle):
    img = synth_image(640, 480, 2)
    e = oracle.extractor(1000, 1.2, 8, 20, 7)
    e(img)
    prev = img
    for l in range(8):
        lv = e.level(l)
        if l > 0:
            prev = cv2.resize(prev, lv.shape[::-1], interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(lv, prev)
        b = e.level(l, blurred=True)
        assert np.array_equal(b, cv2.GaussianBlur(lv, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101))


def test_extractor_candidates_match_per_cell_cv2(oracle):
    """Whole E3 stage against cv2: per-cell FAST with the ini->min fallback (src/ORBextractor.cc:1089-1157)."""
    img = synth_image(640, 480, 4)
    e = oracle.extractor(1000, 1.2, 8, 20, 7)
    e(img)
    det20, det7 = cv2.FastFeatureDetector_create(20, True), cv2.FastFeatureDetector_create(7, True)
    for l in (0, 3, 7):
        lv = e.level(l)
        h, w = lv.shape
        minB, maxBX, maxBY = 16, w - 16, h - 16
        width, height = float(maxBX - minB), float(maxBY - minB)
        nCols, nRows = int(width / 30), int(height / 30)
        wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
        ref = []
        for i in range(nRows):
            iniY = minB + i * hCell
            maxY = min(iniY + hCell + 6, maxBY)
            if iniY >= maxBY - 3:
                continue
            for j in range(nCols):
                iniX = minB + j * wCell
                maxX = min(iniX + wCell + 6, maxBX)
                if iniX >= maxBX - 6:
                    continue
                roi = np.ascontiguousarray(lv[iniY:maxY, iniX:maxX])
                kps = det20.detect(roi)
                if not kps:
                    kps = det7.detect(roi)
                ref += [(k.pt[0] + j * wCell, k.pt[1] + i * hCell, k.response) for k in kps]
        c = e.candidates(l)
        assert [(float(a), float(b), float(r)) for a, b, r in zip(c["x"], c["y"], c["response"])] == ref
