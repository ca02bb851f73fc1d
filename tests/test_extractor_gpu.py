"""GPU parity of the extractor against the CPU oracle, stage by stage and end to end (bit-exact)."""
import numpy as np
import pytest

from synth import synth_image, synth_stereo

pytestmark = pytest.mark.gpu

CONFIGS = [(640, 480, 1000), (1241, 376, 2000)]


def _cmp_kps(a, b):
    assert len(a) == len(b)
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(a[f], b[f]), f


@pytest.mark.parametrize("w,h,nf", CONFIGS)
def test_stages_and_end_to_end(pkg, oracle, w, h, nf):
    img = synth_image(w, h, 3)
    ex = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_width=1280, max_height=512 if h < 512 else 1024)
    oe = oracle.extractor(nf, 1.2, 8, 20, 7)
    kps, desc = ex(img)
    okps, odesc = oe(img)
    (t, nfl, _) = oe.tables()
    assert np.array_equal(ex.mvScaleFactor, t[0]) and np.array_equal(ex.mvInvScaleFactor, t[1])
    assert np.array_equal(ex.mvLevelSigma2, t[2]) and np.array_equal(ex.mvInvLevelSigma2, t[3])
    assert np.array_equal(ex.mnFeaturesPerLevel, nfl)
    for l in range(8):
        assert np.array_equal(ex.debug_level(0, l), oe.level(l)), "pyramid level %d" % l
        ob = oe.level(l, blurred=True)
        if ob is not None:
            assert np.array_equal(ex.debug_level(0, l, blurred=True), ob), "blurred level %d" % l
        xy, resp = ex.debug_candidates(0, l)
        oc = oe.candidates(l)
        assert len(xy) == len(oc), "level %d candidates %d vs %d" % (l, len(xy), len(oc))
        assert np.array_equal(xy[:, 0], oc["x"].astype(np.int32)) and np.array_equal(xy[:, 1], oc["y"].astype(np.int32))
        assert np.array_equal(resp, oc["response"].astype(np.int32))
    _cmp_kps(kps, okps)
    assert np.array_equal(desc, odesc)
    assert len(kps) >= nf


def test_stereo_pair_batch(pkg, oracle):
    w, h, nf = 1241, 376, 2000
    left, right = synth_stereo(w, h, 11)
    ex = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_width=1280, max_height=512, max_batch=2)
    oe = oracle.extractor(nf, 1.2, 8, 20, 7)
    res = ex.extract_batch([left, right])
    for im, (kps, desc) in zip((left, right), res):
        okps, odesc = oe(im)
        _cmp_kps(kps, okps)
        assert np.array_equal(desc, odesc)


def test_low_texture_and_empty(pkg, oracle):
    ex = pkg.ORBextractor(500, 1.2, 8, 20, 7, max_width=640, max_height=480)
    oe = oracle.extractor(500, 1.2, 8, 20, 7)
    # empty image -> silent empty result (src/ORBextractor.cc:1553)
    k, d = ex(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    # flat image: no corners at all
    flat = np.full((480, 640), 77, np.uint8)
    k, d = ex(flat)
    assert len(k) == 0
    # weak texture only: exercises the minThFAST fallback in every cell
    rng = np.random.RandomState(5)
    weak = (100 + rng.randint(0, 24, size=(480, 640))).astype(np.uint8)
    k, d = ex(weak)
    ok, od = oe(weak)
    _cmp_kps(k, ok)
    assert np.array_equal(d, od)


def test_sincosf_device_matches_glibc(pkg, oracle):
    import ctypes
    L = pkg.lib()
    L.b2s_debug_sincosf.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    hi = np.float32(6.2832).view(np.uint32)
    chunk = 1 << 26
    bad = 0
    for start in range(0, int(hi) + 1, chunk):
        bits = np.arange(start, min(start + chunk, int(hi) + 1), dtype=np.uint32)
        x = bits.view(np.float32)
        s = np.zeros_like(x)
        c = np.zeros_like(x)
        rc = L.b2s_debug_sincosf(x.ctypes.data, len(x), s.ctypes.data, c.ctypes.data)
        assert rc == 0
        os_, oc = oracle.sincosf(x)
        bad += int((s.view(np.uint32) != os_.view(np.uint32)).sum()) + int((c.view(np.uint32) != oc.view(np.uint32)).sum())
    assert bad == 0


def test_against_committed_golden(pkg):
    """CUDA path vs tests/golden (no oracle, no /root/reference at run time)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "extract_640x480_seed3.npz"))
    ex = pkg.ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480)
    k, d = ex(synth_image(640, 480, 3))
    assert k.tobytes() == g["kps"].tobytes() and np.array_equal(d, g["desc"])


def test_batch_upload_paths(pkg, oracle):
    """Host-buffer batch entry point: >= 16 images take the chunked two-stream path; images in one contiguous block are
    uploaded with one 1-D copy per chunk and re-pitched on the device (odd width x height: unaligned image starts),
    separately allocated images take one copy each.  All of it must equal the per-image oracle result."""
    w, h, nf, B = 321, 243, 500, 18
    imgs = [synth_image(w, h, 40 + i) for i in range(B)]
    ex = pkg.ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    res_sep = ex.extract_batch(imgs)
    block = np.stack(imgs)
    res_blk = ex.extract_batch([block[i] for i in range(B)])
    oex = oracle.extractor(nf, 1.2, 8, 20, 7)
    for i in range(B):
        ok, od = oex(imgs[i])
        for res in (res_sep, res_blk):
            k, d = res[i]
            assert len(k) == len(ok), i
            for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
                assert np.array_equal(k[f], ok[f]), (i, f)
            assert np.array_equal(d, od), i


def test_chunked_batches_from_two_threads(pkg):
    """b2s_extract_batch cuts batches of >= 16 images into chunks that run on two streams; two handles on two host threads
    overlap them further.  Every chunk owns its slice of the minThFAST fallback list (weak-texture images put EVERY cell on
    that list), so the results must equal the one-image-at-a-time results whatever the interleaving."""
    import threading
    w, h, B = 640, 480, 32
    rng = np.random.RandomState(11)
    imgs = []
    for b in range(B):
        if b % 3 == 0:
            imgs.append((100 + rng.randint(0, 24, size=(h, w))).astype(np.uint8))  # weak texture: fallback in every cell
        else:
            imgs.append(synth_image(w, h, 300 + b))
    one = pkg.ORBextractor(500, 1.2, 8, 20, 7, max_width=w, max_height=h)
    ref = [one(im) for im in imgs]
    exs = [pkg.ORBextractor(500, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B) for _ in range(2)]
    bad = []

    def run(ex, order):
        for it in range(6):
            out = ex.extract_batch([imgs[i] for i in order])
            for j, i in enumerate(order):
                if not (np.array_equal(out[j][0], ref[i][0]) and np.array_equal(out[j][1], ref[i][1])):
                    bad.append((it, i))
    th = [threading.Thread(target=run, args=(exs[0], list(range(B)))), threading.Thread(target=run, args=(exs[1], list(range(B - 1, -1, -1))))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not bad, bad[:8]
