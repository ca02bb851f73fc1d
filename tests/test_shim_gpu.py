"""The C++ shim classes with the reference's signatures (self_commit_orb-slam2_b200/host) against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from synth import synth_image

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_extractor_shim(pkg, oracle, tmp_path):
    host = os.path.join(ROOT, "self_commit_orb-slam2_b200", "host")
    exe = str(tmp_path / "test_shim")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_shim.cc"),
                           os.path.join(host, "ORBextractor.cc"), os.path.join(host, "ORBmatcher.cc"), pkg.LIB_PATH,
                           "-Wl,-rpath," + os.path.dirname(pkg.LIB_PATH)])
    img = synth_image(640, 480, 6)
    raw = str(tmp_path / "img.raw")
    img.tofile(raw)
    out = str(tmp_path / "out.bin")
    subprocess.check_call([exe, raw, "640", "480", out])
    buf = open(out, "rb").read()
    n = int(np.frombuffer(buf[:4], np.int32)[0])
    kps = np.frombuffer(buf[4:4 + 28 * n], pkg.keypoint_dtype)
    desc = np.frombuffer(buf[4 + 28 * n:], np.uint8).reshape(n, 32)
    okps, odesc = oracle.extractor(1000, 1.2, 8, 20, 7)(img)
    assert n == len(okps)
    for f in ("x", "y", "size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(kps[f], okps[f]), f
    assert np.array_equal(desc, odesc)
