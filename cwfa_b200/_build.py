"""In-tree build of the C-ABI CUDA library (nvcc, sm_100a only)."""
import contextlib
import fcntl
import glob
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcwfa_b200.so")
HASH_PATH = LIB_PATH + ".sha256"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-DCWFA_B200",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cwfa_b200 needs the CUDA toolkit to build its kernels")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + sorted(glob.glob(os.path.join(CSRC, "*.h"))) + \
        sorted(glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h")))


def source_hash() -> str:
    """sha256 over the flags and the contents of every source / header the library is built from.  Stored next to the
    .so: staleness does not depend on file mtimes (a snapshot copy to another box does not preserve them)."""
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def built_hash() -> str:
    try:
        with open(HASH_PATH) as f:
            return f.read().strip()
    except OSError:
        return ""


def _stale() -> bool:
    return not os.path.exists(LIB_PATH) or built_hash() != source_hash()


@contextlib.contextmanager
def _build_lock():
    """Inter-process lock: under torchrun every rank imports the package at once; exactly one compiles, the others wait and
    then find a fresh library."""
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ into lib/libcwfa_b200.so (object per file, then link).  Objects and the library are
    written under temporary names and renamed into place, so a concurrent reader never maps a half-written file."""
    if not force and not _stale():
        return LIB_PATH
    with _build_lock():
        if not force and not _stale():          # another process built it while we waited
            return LIB_PATH
        want = source_hash()
        nvcc = _nvcc()
        tag = f".tmp{os.getpid()}"
        objs, procs = [], []
        for src in sources():
            obj = os.path.join(LIB_DIR, os.path.basename(src)[:-3] + ".o")
            objs.append(obj)
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj + tag]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        failed = None
        for cmd, p in procs:
            out, _ = p.communicate()
            if p.returncode != 0 and failed is None:
                failed = "nvcc failed: " + " ".join(cmd) + "\n" + out
            if verbose and out:
                print(out)
        if failed:
            for o in objs:
                with contextlib.suppress(OSError):
                    os.remove(o + tag)
            raise RuntimeError(failed)
        for o in objs:
            os.replace(o + tag, o)
        link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH + tag] + objs
        r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed: " + " ".join(link) + "\n" + r.stdout)
        os.replace(LIB_PATH + tag, LIB_PATH)
        with open(HASH_PATH + tag, "w") as f:
            f.write(want)
        os.replace(HASH_PATH + tag, HASH_PATH)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
