"""Builds libadmpc_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension machinery)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libadmpc_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-ccbin", "/usr/bin/g++", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "admpc.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = src[:-3] + ".o"
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on " + src)
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj


def build(force=False, verbose=False):
    if not (force or stale()):
        return SO
    # headers newer than an object invalidate every object; otherwise only the sources that changed are recompiled
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "admpc.h")]
    th = max(os.path.getmtime(h) for h in hdrs)
    todo, objs = [], []
    for src in sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(th, os.path.getmtime(src)):
            todo.append(src)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(lambda s: _compile(s, verbose), todo))
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++", "-o", SO] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(SO)
