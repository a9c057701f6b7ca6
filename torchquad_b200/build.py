"""In-tree build of libtqb200.so with nvcc for sm_100a (no torch headers, plain C ABI).

`python -m torchquad_b200.build` or `torchquad_b200.build.build()`.  nvcc cross-compiles without a GPU.
Objects go to torchquad_b200/csrc/build/ (git-ignored); the shared library stays next to the package so
it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "libtqb200.so")
SOURCES = ["runtime.cu", "rng_mc.cu", "vegas_map.cu", "vegas_strat.cu", "newton_cotes.cu", "fused.cu", "vegas_unfused.cu", "vegas_small.cu",
           "vegas_driver.cu"]
NVCC_FLAGS = [
    *os.environ.get("TQ_NVCC_EXTRA", "").split(),
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; cannot build libtqb200.so")
    return cand


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link libtqb200.so (skipped when up to date)."""
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "tqb200.h"))
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    stamp = os.path.join(OBJ, "stamp.txt")
    digest = _digest(headers + srcs)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    objs = [o for o, _ in results]
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
