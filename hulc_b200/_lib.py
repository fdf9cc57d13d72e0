"""ctypes binding of `libhulc_b200.so` (C ABI declared in include/hulc_b200.h).

The prototypes are parsed from the header itself, so the binding cannot drift from the declared ABI.  There is no CPU
fallback: if the nvcc-built library is missing this module raises at first use (run `python -c "import
__graft_entry__ as g; g.build()"`).
"""
from __future__ import annotations

import ctypes
import re
from pathlib import Path
from typing import Dict, List, Tuple

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
HEADER = ROOT / "include" / "hulc_b200.h"
LIB_PATH = PKG / "lib" / "libhulc_b200.so"

_CTYPES = {
    "int": ctypes.c_int,
    "unsigned": ctypes.c_uint,
    "unsigned int": ctypes.c_uint,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
}


def parse_header(path: Path = HEADER) -> Dict[str, List[Tuple[str, str]]]:
    """-> {function name: [(c type, parameter name), ...]} for every `int hulc_*(...)` prototype in the header."""
    text = re.sub(r"/\*.*?\*/", "", path.read_text(), flags=re.S)
    protos: Dict[str, List[Tuple[str, str]]] = {}
    for m in re.finditer(r"\bint\s+(hulc_\w+)\s*\(([^)]*)\)\s*;", text):
        params = []
        for raw in m.group(2).split(","):
            raw = " ".join(raw.split())
            if not raw or raw == "void":
                continue
            ty, name = raw.rsplit(" ", 1)
            while name.startswith("*"):
                ty, name = ty + "*", name[1:]
            params.append((ty.replace(" *", "*"), name))
        protos[m.group(1)] = params
    return protos


def _ctype(ty: str):
    if ty.endswith("*"):
        return ctypes.c_void_p
    return _CTYPES[ty.replace("const ", "").strip()]


class HulcError(RuntimeError):
    pass


class Library:
    """Typed handle on the shared library; `lib.hulc_gemm(...)` raises HulcError on a non-zero return."""

    def __init__(self, path: Path, allow_missing: bool = False):
        if not Path(path).exists():
            raise HulcError(
                f"{path} not found: the hulc_b200 CUDA library has not been built. Build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` — there is no CPU fallback."
            )
        self.path = Path(path)
        self.cdll = ctypes.CDLL(str(path))
        self.protos = parse_header()
        self.missing = []
        for name, params in self.protos.items():
            try:
                fn = getattr(self.cdll, name)  # AttributeError if the library does not export a declared symbol
            except AttributeError:
                if not allow_missing:
                    raise
                self.missing.append(name)
                setattr(self, name, self._unavailable(name))
                continue
            fn.restype = ctypes.c_int
            fn.argtypes = [_ctype(t) for t, _ in params]
            setattr(self, name, self._wrap(name, fn))

    @staticmethod
    def _unavailable(name):
        def call(*args):
            raise HulcError(f"{name} is not part of this build of the kernel sources (tcgen05 kernels only exist in the nvcc build)")

        return call

    @staticmethod
    def _wrap(name, fn):
        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise HulcError(f"{name} failed with CUDA error {rc}")

        call.__name__ = name
        return call


_LIB = None


def lib() -> Library:
    global _LIB
    if _LIB is None:
        _LIB = Library(LIB_PATH)
    return _LIB
