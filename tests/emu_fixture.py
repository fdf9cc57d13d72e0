"""pytest fixture that points `hulc_b200.ops` at the host-emulated build of the kernel sources (tests/emu).  CPU tests
only — the `-m gpu` tests never use it."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "emu"))


@pytest.fixture(scope="session")
def emu_lib_path():
    import build_emu

    return build_emu.build()


@pytest.fixture()
def emu(emu_lib_path, monkeypatch):
    from hulc_b200 import _lib, ops

    monkeypatch.setattr(_lib, "_LIB", _lib.Library(emu_lib_path))
    monkeypatch.setattr(ops, "_DEVICE_TYPE", "cpu")
    monkeypatch.setattr(ops, "_workspaces", {})
    return ops
