"""Build libfmi_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

The library is linked against the static CUDA runtime only (the driver entry points it needs
for TMA descriptors are resolved at run time), so it loads on a CPU-only box; every compute
entry point then fails loudly with a CUDA error.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
ROOT = PKG.parent
BUILD = ROOT / "build" / "fmi_b200"
LIB = PKG / "libfmi_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(ROOT / "include"),
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; cannot build libfmi_b200.so")


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest(src: Path) -> str:
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))):
        h.update(hdr.read_bytes())
    return h.hexdigest()


def _compile_one(nvcc: str, src: Path, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".sha1")
    dig = _digest(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == dig:
        return obj
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (BUILD / (src.stem + ".ptxas.log")).write_text(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(f"[fmi_b200.build] compiled {src.name}\n")
    stamp.write_text(dig)
    return obj


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


DIGEST = LIB.with_name(LIB.name + ".digest")


def source_digest() -> str:
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).replace(str(ROOT), "").encode())
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((ROOT / "include").glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()


def stale() -> bool:
    """True when the linked library was not built from the sources that are here now: the digest of csrc/ + include/ written
    next to it at link time differs (content, not mtime: a copy of the tree to another box must not look stale)."""
    if not LIB.exists() or not DIGEST.exists():
        return True
    return DIGEST.read_text().strip() != source_digest()


def build(force: bool = False, verbose: bool = True) -> Path:
    """Compile every csrc/*.cu for sm_100a and link face_mask_inpaint_b200/libfmi_b200.so — under an exclusive file lock (N ranks
    of a torchrun job on a fresh checkout must not run nvcc into the same objects), linking to a temporary name that is renamed
    into place, so that no process can dlopen a partially written library."""
    import fcntl
    BUILD.mkdir(parents=True, exist_ok=True)
    with open(BUILD / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and LIB.exists() and not stale():
                return LIB            # another process built it while this one waited for the lock
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> Path:
    nvcc = _nvcc()
    if force:
        for f in BUILD.glob("*.sha1"):
            f.unlink()
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or stale() or LIB.stat().st_mtime < newest:
        tmp = LIB.with_name(f".{LIB.name}.{os.getpid()}.tmp")
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
               "-cudart", "static", "-o", str(tmp), *map(str, objs)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            tmp.unlink(missing_ok=True)
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        os.replace(tmp, LIB)
        DIGEST.write_text(source_digest())
        if verbose:
            sys.stderr.write(f"[fmi_b200.build] linked {LIB}\n")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
