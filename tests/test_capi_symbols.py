"""libb200slam.so loads (no GPU needed) and exports every entry point include/b200slam.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "b200slam.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(pkg):
    names = _declared()
    assert len(names) >= 20
    L = ctypes.CDLL(pkg.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the product fails loudly (B2S_ERR_NO_DEVICE); it never routes through the oracle."""
    if pkg.device_count() > 0:
        return  # running on a GPU box: covered by the gpu tests
    import pytest
    with pytest.raises(pkg.B200SlamError) as ei:
        pkg.ORBextractor(1000, 1.2, 8, 20, 7)
    assert ei.value.code == pkg.ERR_NO_DEVICE
    with pytest.raises(pkg.B200SlamError):
        pkg.ORBmatcher()
    with pytest.raises(pkg.B200SlamError):
        pkg.Optimizer()


def test_product_does_not_link_oracle(pkg):
    import subprocess
    out = subprocess.check_output(["ldd", pkg.LIB_PATH]).decode()
    assert "orb_oracle" not in out
    syms = subprocess.check_output(["nm", "-D", pkg.LIB_PATH]).decode()
    assert "orc_" not in syms
