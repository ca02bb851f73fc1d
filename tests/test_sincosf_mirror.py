"""The device restates glibc's sincosf kernels (the reference calls cos/sin(float) at src/ORBextractor.cc:181).
This CPU test checks the same operation sequence (written in C with explicit fma) against this box's libm for EVERY
float in [0, 2*pi]; the GPU test test_sincosf_device_matches_glibc checks the device code against libm."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_mirror_matches_libm_exhaustively(tmp_path):
    exe = str(tmp_path / "sincosf_mirror")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-o", exe,
                           os.path.join(HERE, "sincosf_mirror.c"), "-lm"])
    out = subprocess.check_output([exe], timeout=600).decode()
    assert "sin mismatches 0 cos mismatches 0" in out, out
