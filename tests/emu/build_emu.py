"""Build the host-emulated kernel library (development aid; see cuda_emu.h).  Output: tests/emu/_build/libhulc_b200_emu.so"""
import hashlib
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "hulc_b200" / "csrc"
OUT = HERE / "_build"


def build(verbose=False) -> Path:
    """(Serialised with a file lock: the test workers of a parallel run all ask for the library at start-up.)"""
    import fcntl

    OUT.mkdir(exist_ok=True)
    with open(OUT / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        return _build(verbose)


def _build(verbose=False) -> Path:
    OUT.mkdir(exist_ok=True)
    srcs = sorted(p for p in CSRC.glob("*.cu") if "_tc" not in p.stem and "_tma" not in p.stem) + [HERE / "cuda_emu.cc"]  # tcgen05 / TMA kernels are not emulated
    deps = srcs + sorted(CSRC.glob("*.cuh")) + [HERE / "cuda_emu.h"]
    hdr_hash = hashlib.sha1(b"".join(p.read_bytes() for p in deps if p.suffix in (".cuh", ".h"))).hexdigest()[:12]
    objs = []
    procs = []
    for s in srcs:
        h = hashlib.sha1(s.read_bytes()).hexdigest()[:12]
        o = OUT / f"{s.stem}.{h}.{hdr_hash}.o"
        objs.append(o)
        if o.exists():
            continue
        for old in OUT.glob(f"{s.stem}.*.o"):
            old.unlink()
        cmd = ["g++", "-O2", "-g0", "-fPIC", "-std=c++17", "-fno-strict-aliasing", "-Wno-attributes", "-Wno-unknown-pragmas",
               "-DHULC_HOST_EMULATION", "-include", str(HERE / "cuda_emu.h"), "-I", str(CSRC), "-I", str(ROOT / "include"),
               "-x", "c++", "-c", str(s), "-o", str(o)]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode())
            raise RuntimeError(f"emu build failed for {s}")
    lib = OUT / "libhulc_b200_emu.so"
    if procs or not lib.exists():
        subprocess.check_call(["g++", "-shared", "-o", str(lib)] + [str(o) for o in objs] + ["-lpthread"])
    return lib


if __name__ == "__main__":
    print(build(verbose=True))
