"""Build libgdn_b200.so in-tree with nvcc for sm_100a (no torch involved).

    python -m gdn_b200.build            # incremental
    python -m gdn_b200.build --force
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgdn_b200.so")
SOURCES = ["api.cu", "attention.cu", "dwide.cu", "graph_build.cu", "scoring.cu", "csr.cu", "windows.cu", "metrics.cu", "optim.cu"]
HOST_SOURCES = ["host_stage.cpp"]        # host-only C++ (g++): the staging copy of the feed
CXX = os.environ.get("CXX", "g++")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _newer(src, dst):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "gdn_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(name, force):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name.replace(".cu", ".o"))
    if not force and not _newer(src, obj):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    with open(obj + ".ptxas.log", "w") as f:
        f.write(r.stderr)
    return obj, r.stderr


def _compile_host(name, force):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name.replace(".cpp", ".o"))
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src):
        return obj, ""
    cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-pthread", "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{CXX} failed for {name}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    extra = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu") and f not in srcs)
    srcs += extra
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    results += [_compile_host(h, force) for h in HOST_SOURCES if os.path.exists(os.path.join(CSRC, h))]
    objs = [o for o, _ in results]
    relink = force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if relink:
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for _, log in results:
            if log:
                print(log)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
