"""In-tree build of librtsds_b200.so (sm_100a only) with plain nvcc.

No torch headers are involved: the library is a true C ABI (include/rtsds_b200.h)
and is loaded with ctypes (rtsds_b200/_lib.py).  Objects are cached under
rtsds_b200/build/ keyed by a hash of the source, the headers and the flags.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "build"
LIB = PKG / "librtsds_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; librtsds_b200 cannot be built")
    return cand


def _digest(src: Path, headers: list[Path], extra: list[str]) -> str:
    h = hashlib.sha256()
    h.update(src.read_bytes())
    for p in headers:
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS + extra).encode())
    return h.hexdigest()[:16]


def build(verbose: bool = False, force: bool = False, defines: list[str] | None = None) -> Path:
    """Compile every csrc/*.cu and link librtsds_b200.so. Returns its path.  Serialised across processes by a file lock
    (under torchrun every rank may find the library missing at once)."""
    import fcntl

    BUILD.mkdir(exist_ok=True)
    with open(BUILD / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(verbose, force, defines)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def source_digest() -> str:
    """Digest of every source / header / flag that goes into the library (stale-library check in _lib.lib())."""
    h = hashlib.sha256()
    for p in sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h")):
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _build_locked(verbose: bool, force: bool, defines: list[str] | None) -> Path:
    extra = [f"-D{d}" for d in (defines or [])]
    nvcc = _nvcc()
    BUILD.mkdir(exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted((PKG.parent / "include").glob("*.h"))
    sources = sorted(CSRC.glob("*.cu"))
    objs: list[Path] = []
    jobs = []
    for src in sources:
        obj = BUILD / f"{src.stem}.{_digest(src, headers, extra)}.o"
        objs.append(obj)
        if force or not obj.exists():
            for old in BUILD.glob(f"{src.stem}.*.o"):
                old.unlink()
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = BUILD / f"{src.stem}.ptxas.log"
        log.write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(f"[build] {src.name} ok", file=sys.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(compile_one, jobs))
    if jobs or force or not LIB.exists():
        # link to a private name and rename into place: a concurrent loader never sees a half-written library
        tmp = LIB.with_name(f".{LIB.name}.{os.getpid()}.tmp")
        cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", str(tmp), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)
        if verbose:
            print(f"[build] linked {LIB}", file=sys.stderr)
    (BUILD / "lib.digest").write_text(source_digest())
    return LIB


if __name__ == "__main__":
    build(verbose=True, force="--force" in sys.argv)
    print(LIB)
