"""Build the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m synthanatomy_b200.build          # -> synthanatomy_b200/lib/libsynthanatomy_b200.so
"""
from __future__ import annotations

import concurrent.futures as cf
import contextlib
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIBNAME = "libsynthanatomy_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; synthanatomy_b200 has no non-CUDA fallback")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def library_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def stamp_path() -> str:
    """digest of the sources the .so was built from; lives next to the .so so that it travels with it"""
    return library_path() + ".digest"


def source_digest() -> str:
    root = os.path.dirname(HERE)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers += [os.path.join(root, "include", f) for f in sorted(os.listdir(os.path.join(root, "include")))
                if f.endswith(".h")]
    srcs = [os.path.join(CSRC, f) for f in _sources()]
    # paths enter the digest relative to the repo root: the same tree hashes the same wherever it is checked out
    h = hashlib.sha1()
    for p in sorted(srcs + headers):
        with open(p, "rb") as f:
            h.update(os.path.relpath(p, root).encode()); h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    out, stamp = library_path(), stamp_path()
    return os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == source_digest()


@contextlib.contextmanager
def _build_lock():
    """one builder at a time across processes (every rank of a torchrun job calls load() at once)"""
    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lf:
        fcntl.flock(lf, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(lf, fcntl.LOCK_UN)


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    out = library_path()
    if not force and is_current():
        return out
    with _build_lock():
        if not force and is_current():          # another process built it while this one waited for the lock
            return out
        nvcc = _nvcc()
        srcs = [os.path.join(CSRC, f) for f in _sources()]
        digest = source_digest()
        tag = f".{os.getpid()}.tmp"

        def compile_one(src):
            obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
            cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj + tag]
            if verbose:
                cmd.insert(1, "-Xptxas"); cmd.insert(2, "-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
            os.replace(obj + tag, obj)
            return obj

        with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
            objs = list(ex.map(compile_one, srcs))
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out + tag, *objs]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(out + tag, out)              # a concurrent loader sees the old library or the new one, never half
        with open(stamp_path() + tag, "w") as f:
            f.write(digest)
        os.replace(stamp_path() + tag, stamp_path())
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
