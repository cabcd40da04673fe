"""Builds libstarst3r_b200.so (sm_100a only) in-tree with nvcc.

Usage: python -m starst3r_b200.build [--force] [--verbose]
The shared library exposes only the C ABI of include/starst3r_b200.h; there is
no torch / pybind dependency, Python binds it with ctypes (starst3r_b200/_lib.py).
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libstarst3r_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# Files whose arithmetic must not be FMA-contracted (bit-exact tile/bin indices).
NO_FMAD = {"gs_project.cu", "gs_backward.cu"}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path, flags):
    h = hashlib.sha256()
    h.update(" ".join(flags).encode())
    deps = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    deps += [os.path.join(os.path.dirname(HERE), "include", "starst3r_b200.h"), path]
    for p in deps:
        with open(p, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(src, force, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src[:-3] + ".o")
    flags = ARCH + COMMON + (["-fmad=false"] if src in NO_FMAD else [])
    stamp = obj + ".sha"
    dig = _digest(path, flags)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return src, "", False
    cmd = [NVCC] + flags + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    return src, r.stderr, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    rebuilt = any(r[2] for r in results)
    if verbose:
        for src, log, did in results:
            if did:
                print(f"== {src}\n{log}")
    if rebuilt or not os.path.exists(LIB):
        objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined",
                                                               "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(lib)
