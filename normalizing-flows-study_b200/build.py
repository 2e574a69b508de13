"""Build libnfb200.so (sm_100a only) in-tree with nvcc.

    python normalizing-flows-study_b200/build.py [--force]

Every translation unit under csrc/ is compiled in parallel
(`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3`) and linked into
`normalizing-flows-study_b200/lib/libnfb200.so`.  nvcc cross-compiles without a GPU.
No torch headers are involved: the library is a plain C-ABI shared object (include/nfb200.h).
"""
import concurrent.futures as cf
import glob
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libnfb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src, headers_digest, force):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp = obj + ".stamp"
    want = headers_digest + _digest([src])
    if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == want:
        return obj, 0.0
    import time
    t = time.time()
    cmd = [NVCC] + ARCH + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return obj, time.time() - t


def build(force=False, verbose=True):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(PKG, "..", "include", "*.h"))
    hd = _digest(headers)
    # the object directory does not travel to the GPU box (.gpurunignore) but the library does: a digest of all
    # sources stored beside it says whether the shipped library is current
    whole = _digest(headers + srcs) + " ".join(ARCH + FLAGS)
    lib_stamp = LIB + ".digest"
    if not force and os.path.exists(LIB) and os.path.exists(lib_stamp) and open(lib_stamp).read() == whole:
        if verbose:
            print(f"  up to date: {LIB}", flush=True)
        return LIB
    objs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for obj, dt in ex.map(lambda s: _compile(s, hd, force), srcs):
            objs.append(obj)
            if verbose and dt > 0:
                print(f"  nvcc {os.path.basename(obj)} {dt:.1f}s", flush=True)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        # link beside the target and rename: a snapshot of the tree (gpurun) never sees a half-written library
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB + ".tmp"] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(LIB + ".tmp", LIB)
        if verbose:
            print(f"  linked {LIB}", flush=True)
    with open(lib_stamp, "w") as f:
        f.write(whole)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
