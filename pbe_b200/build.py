"""In-tree build of the sm_100a C-ABI library (``pbe_b200/lib/libpbe_b200.so``).

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
OBJDIR = ROOT / "build"
LIB = LIBDIR / "libpbe_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + list((ROOT.parent / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def tree_digest() -> str:
    """One digest over the compile flags and every source / header the library is built from.  It is written next to the
    library at link time (``lib/libpbe_b200.sha``, which travels with the .so) and compared at load time, so a library
    that no longer matches the sources is rebuilt (or refused) instead of being loaded with stale signatures."""
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) +
                    list((ROOT.parent / "include").glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


LIB_STAMP = LIBDIR / "libpbe_b200.sha"


def library_is_current() -> bool:
    return LIB.exists() and LIB_STAMP.exists() and LIB_STAMP.read_text().strip() == tree_digest()


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile every ``csrc/*.cu`` for sm_100a and link the shared library. Incremental by content hash; serialised across
    processes by a file lock (several ranks importing at once build once)."""
    import fcntl
    LIBDIR.mkdir(exist_ok=True)
    with open(LIBDIR / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and library_is_current():
                return LIB
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    nvcc = _nvcc()
    OBJDIR.mkdir(exist_ok=True)
    srcs = _sources()
    jobs = []
    for src in srcs:
        obj = OBJDIR / (src.stem + ".o")
        stamp = OBJDIR / (src.stem + ".sha")
        dig = _digest(src)
        if force or not obj.exists() or not stamp.exists() or stamp.read_text() != dig:
            jobs.append((src, obj, stamp, dig))

    def compile_one(job):
        src, obj, stamp, dig = job
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        (OBJDIR / (src.stem + ".log")).write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        stamp.write_text(dig)
        if verbose:
            print(r.stderr)
        return src.name

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or not LIB.exists():
        objs = [str(OBJDIR / (s.stem + ".o")) for s in srcs]
        cmd = [nvcc, "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    LIB_STAMP.write_text(tree_digest())
    return LIB


if __name__ == "__main__":
    import sys
    print(build_native(force="--force" in sys.argv, verbose=True))
