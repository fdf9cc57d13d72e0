"""pytest fixtures selecting the build of the kernel sources a test runs against.

`emu`  — the host SIMT emulator build (tests/emu), CPU tensors; used by the `-m "not gpu"` suite.
`K`    — parametrised: "emu" (as above) and "cuda" (the nvcc-built libhulc_b200.so on the B200, marked `gpu`): the same
         kernel test bodies run on both; the cuda variant moves every tensor argument to the device (preserving strides
         and storage sharing) and copies the storages back after each call, so the bodies stay device-agnostic."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "emu"))


@pytest.fixture(scope="session")
def emu_lib_path():
    import build_emu

    return build_emu.build()


@pytest.fixture()
def emu(emu_lib_path, monkeypatch):
    from hulc_b200 import _lib, ops

    monkeypatch.setattr(_lib, "_LIB", _lib.Library(emu_lib_path, allow_missing=True))  # the tcgen05 kernels are not emulated
    monkeypatch.setattr(ops, "_DEVICE_TYPE", "cpu")
    monkeypatch.setattr(ops, "_workspaces", {})
    return ops


class DeviceProxy:
    """Calls hulc_b200.ops functions with CPU tensors by mirroring their storages on the GPU for the duration of a call."""

    def __init__(self, ops_module, device="cuda"):
        self._ops, self._device = ops_module, device
        self.NO_DROP = ops_module.NO_DROP

    def Drop(self, p=0.0, seed=0, site=0, keep=None):
        d = self._ops.Drop(p, seed, site, keep)
        return d

    def __getattr__(self, name):
        fn = getattr(self._ops, name)

        def call(*args, **kw):
            mirrors = {}

            def conv(a):
                if torch.is_tensor(a) and a.device.type == "cpu":
                    st = a.untyped_storage()
                    key = st.data_ptr()
                    if key not in mirrors:
                        n = st.nbytes() // a.element_size()
                        flat = torch.as_strided(a, (n,), (1,), 0)
                        mirrors[key] = (flat, flat.to(self._device))
                    return torch.as_strided(mirrors[key][1], a.size(), a.stride(), a.storage_offset())
                if isinstance(a, self._ops.Drop) and a.keep is not None:
                    return self._ops.Drop(a.p, a.seed, a.site, conv(a.keep))
                return a

            res = fn(*[conv(a) for a in args], **{k: conv(v) for k, v in kw.items()})
            torch.cuda.synchronize()
            for flat, dev in mirrors.values():
                back = dev.cpu()
                if not torch.equal(flat.detach().view(torch.uint8), back.view(torch.uint8)):  # read-only arguments keep their autograd version
                    flat.copy_(back)
            if torch.is_tensor(res) and res.device.type != "cpu":
                # results that alias an argument: return the CPU original; fresh results: a CPU copy
                for a in list(args) + list(kw.values()):
                    if torch.is_tensor(a) and a.shape == res.shape and conv(a).data_ptr() == res.data_ptr():
                        return a
                return res.cpu()
            return res

        return call


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def K(request):
    if request.param == "emu":
        return request.getfixturevalue("emu")
    if not torch.cuda.is_available():
        pytest.fail("the gpu-marked kernel tests need a CUDA device (there is no CPU fallback)")
    from hulc_b200 import ops

    return DeviceProxy(ops)
