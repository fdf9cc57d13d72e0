"""Build `hulc_b200/lib/libhulc_b200.so` from `hulc_b200/csrc/*.cu` with nvcc for sm_100a (in-tree, so the library
travels with the repo snapshot to the GPU box).  Objects are cached by content hash under `build/`."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
OBJ = ROOT / "build" / "obj"
LIB = PKG / "lib" / "libhulc_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found")


def build(verbose: bool = False, force: bool = False) -> Path:
    """(Serialised with a file lock: concurrent callers — parallel test workers — wait for the first one's build.)"""
    import fcntl

    OBJ.mkdir(parents=True, exist_ok=True)
    with open(OBJ / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build(verbose, force)


def _build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(parents=True, exist_ok=True)
    LIB.parent.mkdir(parents=True, exist_ok=True)
    srcs = sorted(CSRC.glob("*.cu"))
    hdr_hash = hashlib.sha1(
        b"".join(p.read_bytes() for p in sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h")))
        + " ".join(NVCC_FLAGS).encode()
    ).hexdigest()[:12]
    objs, procs = [], []
    for s in srcs:
        h = hashlib.sha1(s.read_bytes()).hexdigest()[:12]
        o = OBJ / f"{s.stem}.{h}.{hdr_hash}.o"
        objs.append(o)
        if o.exists() and not force:
            continue
        for old in OBJ.glob(f"{s.stem}.*.o"):
            old.unlink()
        cmd = [nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + ["-I", str(CSRC), "-I", str(ROOT / "include"), "-c", str(s), "-o", str(o)]
        if verbose:
            print(" ".join(cmd), flush=True)
        procs.append((s, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for s, o, p in procs:
        out, _ = p.communicate()
        log = out.decode()
        (OBJ / f"{s.stem}.ptxas.log").write_text(log)
        if p.returncode != 0:
            sys.stderr.write(log)
            failed = True
            if o.exists():
                o.unlink()
        elif verbose:
            sys.stdout.write(log)
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not LIB.exists() or force:
        cmd = [nvcc(), "-shared", "-o", str(LIB)] + [str(o) for o in objs] + ["-lcudart", "-lcuda"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
