"""Builds gridpp_b200/libgridpp_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = ["capi.cu", "points.cu", "oi.cu", "neighbourhood.cu", "neighbourhood_tma.cu", "quantile_tma.cu", "ensemble_forms.cu", "thresholds.cu", "ensi.cu", "stats.cu", "gridding.cu", "multi_gpu.cu", "ensi_multi.cu"]
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
    """Compiles every source to an object file in parallel (one nvcc per file), then links the shared library."""
    if not force and not needs_build():
        return LIB
    import concurrent.futures
    import tempfile
    srcs = [os.path.join(HERE, "csrc", s) for s in SOURCES if os.path.exists(os.path.join(HERE, "csrc", s))]
    common = [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-ccbin", "/usr/bin/g++",
              "-Xcompiler", "-fPIC,-fopenmp,-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc")]
    if verbose:
        common.insert(1, "-Xptxas=-v")
    with tempfile.TemporaryDirectory(prefix="gridpp_b200_build_") as tmp:
        def compile_one(src):
            obj = os.path.join(tmp, os.path.basename(src) + ".o")
            res = subprocess.run(common + ["-c", src, "-o", obj], capture_output=True, text=True)
            return src, obj, res
        objs = []
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as pool:
            for src, obj, res in pool.map(compile_one, srcs):
                if res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                    raise RuntimeError("nvcc failed compiling " + os.path.basename(src))
                if verbose:
                    sys.stderr.write(res.stderr)
                objs.append(obj)
        res = subprocess.run([nvcc_path(), "-shared", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lgomp"],
                             capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed linking libgridpp_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
