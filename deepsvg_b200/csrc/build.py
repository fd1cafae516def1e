"""Builds libdsvg_b200.so (in-tree, next to the package) with nvcc for sm_100a.

Usage: python -m deepsvg_b200.csrc.build [--force] [--verbose]
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

CSRC = Path(__file__).resolve().parent
PKG = CSRC.parent
LIB = PKG / "libdsvg_b200.so"
STAMP = PKG / ".libdsvg_b200.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-shared",
]


def sources():
    return sorted(CSRC.glob("*.cu"))


def _digest():
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "dsvg_b200.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Idempotent and safe to call from several processes at once (one rank per GPU under torchrun): an exclusive
    file lock serialises the compilation; the ranks that waited find the stamp up to date and return."""
    import fcntl
    dig = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == dig:
        return LIB
    with open(PKG / ".build.lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(dig, force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(dig, force, verbose):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == dig:
        return LIB
    if not Path(nvcc).exists():
        if LIB.exists():
            return LIB  # GPU box: use the prebuilt library that travelled with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libdsvg_b200.so")
    objs = []
    build_dir = PKG / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in sources():
        obj = build_dir / (src.stem + ".o")
        cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "-shared"] + ["-c", str(src), "-o", str(obj)]
        if verbose:
            cmd += ["-Xptxas", "-v"]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src.name} ---\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {src.name} ---\n{out}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    tmp = LIB.with_suffix(".so.tmp")
    cmd = [nvcc, "-shared", "-o", str(tmp)] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)     # atomic: a concurrent loader never maps a half-written library
    STAMP.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
