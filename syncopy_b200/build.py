"""
In-tree build of libspyb200.so with nvcc for sm_100a (cross-compiles without a GPU).

    python -m syncopy_b200.build [--force] [--verbose]

The shared library lands in syncopy_b200/lib/ (git-ignored, but it travels to the GPU box
with the repository snapshot).  Each .cu is compiled to an object with
`-gencode arch=compute_100a,code=sm_100a -lineinfo` and linked with the static CUDA runtime.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(LIBDIR, "libspyb200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path):
    h = hashlib.sha256()
    # an object depends on its .cu and on every header (cheap, conservative)
    deps = [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    deps.append(os.path.join(os.path.dirname(HERE), "include", "spyb200.h"))
    for d in deps:
        with open(d, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(ARCH + CFLAGS).encode())
    return h.hexdigest()


def _compile(src, force, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJDIR, src[:-3] + ".o")
    stamp = obj + ".sha"
    dig = _digest(path)
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ""
    cmd = [NVCC] + ARCH + CFLAGS + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as fh:
        fh.write(dig)
    with open(obj + ".ptxas.log", "w") as fh:
        fh.write(res.stderr)
    return obj, res.stderr if verbose else ""


def build(force=False, verbose=False):
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [r[0] for r in results]
    for _, log in results:
        if log:
            print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(lib)
