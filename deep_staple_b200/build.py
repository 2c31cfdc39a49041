"""Build libdeep_staple_b200.so (the C-ABI library of include/deep_staple_b200.h) in-tree.

    python -m deep_staple_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The library links cudart statically and has no
torch / Python dependency; the Python package binds it with ctypes (deep_staple_b200/_binding.py).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "_lib")
LIB = os.path.join(LIB_DIR, "libdeep_staple_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")
INCLUDE = os.path.join(ROOT, "include")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC",
]
INCLUDE_FLAGS = ["-I", INCLUDE, "-I", CSRC]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the sm_100a library cannot be built")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))


def _fingerprint() -> str:
    h = hashlib.sha256()
    for path in sorted(glob.glob(os.path.join(CSRC, "*"))) + [os.path.join(INCLUDE, "deep_staple_b200.h"), __file__]:
        with open(path, "rb") as fh:
            h.update(os.path.basename(path).encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp = LIB + ".stamp"
    if not (os.path.exists(LIB) and os.path.exists(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _fingerprint()


def _headers_digest() -> str:
    h = hashlib.sha256()
    for path in sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc"))
                       + [os.path.join(INCLUDE, "deep_staple_b200.h")]):
        with open(path, "rb") as fh:
            h.update(os.path.basename(path).encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, verbose: bool, headers: str = "") -> str:
    """One translation unit -> build/obj/<name>.o; skipped when source, headers and flags are unchanged."""
    obj = os.path.join(OBJ_DIR, os.path.splitext(os.path.basename(src))[0] + ".o")
    with open(src, "rb") as fh:
        key = hashlib.sha256(fh.read() + headers.encode()).hexdigest()
    stamp = obj + ".stamp"
    if headers and not verbose and os.path.exists(obj) and os.path.exists(stamp):
        with open(stamp) as fh:
            if fh.read().strip() == key:
                return obj
    cmd = [_nvcc(), *NVCC_FLAGS, *INCLUDE_FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    with open(stamp, "w") as fh:
        fh.write(key)
    return obj


def build_variant(name: str, extra_flags) -> str:
    """Experiment helper: build the library with extra nvcc flags (e.g. -DDS_LOSS_MIN_BLOCKS=5) into
    build/alt/<name>/; select it at run time with DS_LIB_PATH (see _binding.py)."""
    out_dir = os.path.join(ROOT, "build", "alt", name)
    os.makedirs(out_dir, exist_ok=True)
    objs = []

    def one(src):
        obj = os.path.join(out_dir, os.path.splitext(os.path.basename(src))[0] + ".o")
        res = subprocess.run([_nvcc(), *NVCC_FLAGS, *extra_flags, *INCLUDE_FLAGS, "-c", src, "-o", obj],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(res.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as pool:
        objs = list(pool.map(one, sources()))
    lib = os.path.join(out_dir, "libdeep_staple_b200.so")
    res = subprocess.run([_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stderr)
    for obj in objs:  # variants travel to the GPU box with the snapshot: keep only the library
        os.remove(obj)
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a and link the shared library.  Returns its path."""
    if not force and is_current():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    os.makedirs(LIB_DIR, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as pool:
        headers = "" if force else _headers_digest()
        objs = list(pool.map(lambda s: _compile(s, verbose, headers), srcs))
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(LIB + ".stamp", "w") as fh:
        fh.write(_fingerprint())
    return LIB


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
