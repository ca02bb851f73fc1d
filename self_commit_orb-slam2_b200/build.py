"""Builds libb200slam.so (sm_100a only) in-tree with nvcc, and the CPU oracle with g++.

The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200slam.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liborb_oracle.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler",
              "-fPIC,-O3"]
# extractor/matcher: bit-exact float arithmetic, never contract a*b+c; LocalBA is FP64 with a 1e-5 bar -> FMA allowed
NO_FMAD = {"extractor.cu", "extractor_tile.cu", "matcher.cu", "common.cu"}


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def cuda_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_of(src):
    """Files whose change makes `src`'s object stale (the include graph of csrc/ is small and fixed)."""
    base = os.path.basename(src)
    deps = [src, os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "b200slam.h")]
    if base.startswith("extractor"):
        deps.append(os.path.join(CSRC, "extractor.cuh"))
    if base == "extractor.cu":
        deps.append(os.path.join(ROOT, "data", "orb_pattern_31.inc"))
    return deps


def build_cuda(force=False, verbose=False):
    srcs = cuda_sources()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for s in srcs:
        o = s[:-3] + ".o"
        objs.append(o)
        if not force and not verbose and _newer(o, _deps_of(s)):
            continue  # object up to date (localba.cu alone takes minutes to compile)
        cmd = [nvcc] + NVCC_FLAGS + (["--fmad=false"] if os.path.basename(s) in NO_FMAD else []) + (
            ["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if not procs and not force and _newer(LIB, objs):
        return LIB
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h"))]
    srcs.append(os.path.join(ROOT, "data", "orb_pattern_31.inc"))
    if not force and _newer(ORACLE_LIB, srcs):
        return ORACLE_LIB
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "-B", "liborb_oracle.so"])
    return ORACLE_LIB


def build_oracle_ref():
    """oracle/_ref: the reference's own sources compiled in place against oracle/refshim (only where /root/reference is
    mounted, i.e. in the build container; the GPU box uses the prebuilt files).  Checker infrastructure, never product."""
    ref = os.environ.get("B2S_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "src")):
        return None
    subprocess.check_call(["make", "-s", "-j8", "-C", ORACLE_DIR, "ref", "REF=" + ref])
    return os.path.join(ORACLE_DIR, "_ref")


if __name__ == "__main__":
    build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle(force="--force" in sys.argv)
    build_oracle_ref()
    print("built", LIB, ORACLE_LIB)
