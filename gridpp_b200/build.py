"""Builds gridpp_b200/libgridpp_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = ["capi.cu", "points.cu", "oi.cu", "neighbourhood.cu", "neighbourhood_tma.cu", "quantile_tma.cu", "ensemble_forms.cu", "thresholds.cu", "ensi.cu"]
LIB = os.path.join(HERE, "libgridpp_b200.so")


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    deps.append(os.path.join(ROOT, "include", "gridpp_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    srcs = [os.path.join(HERE, "csrc", s) for s in SOURCES if os.path.exists(os.path.join(HERE, "csrc", s))]
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-fopenmp,-O2", "-shared",
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc"), "-o", LIB] + srcs + ["-lgomp"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libgridpp_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
