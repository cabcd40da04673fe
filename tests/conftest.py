import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def emu_lib(tmp_path_factory):
    """libst3r_emu.so: the C ABI (everything but the tcgen05 matcher) compiled for the host with its kernel launches
    rewritten onto the SIMT emulator (tests/host/build_emu_lib.py), typed from include/starst3r_b200.h.  Built once."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "tests", "host"))
    import build_emu_lib
    from starst3r_b200 import _lib
    path, n_launches = build_emu_lib.build(str(tmp_path_factory.mktemp("emu_lib")))
    assert n_launches >= 50
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in _lib.parse_header().items():
        fn = getattr(lib, name)          # the emulated library exports the complete C ABI (tcgen05 entry points stubbed)
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


@pytest.fixture
def emu_backend(emu_lib, monkeypatch):
    """Runs the product's Python glue on CPU tensors against the emulated library.  The glue refuses anything but CUDA
    by design, so the TEST swaps the library handle, the stream getter and torch.cuda.device for host stand-ins."""
    import contextlib
    import torch
    from starst3r_b200 import _lib
    monkeypatch.setattr(_lib, "load", lambda: emu_lib)
    monkeypatch.setattr(_lib, "stream_ptr", lambda: None)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "require_cuda_device", lambda device, what="": torch.device(device if device is not None else "cpu"))
    monkeypatch.setattr(_lib, "last_error", lambda: emu_lib.st3r_last_error().decode(errors="replace"))
    monkeypatch.setattr(torch.cuda, "device", lambda dev: contextlib.nullcontext())
    yield emu_lib
    assert emu_lib.st3r_emu_launch_failed() == 0, "emulator deadlock / unsupported launch"
