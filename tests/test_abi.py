"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares; argument validation works without a GPU (no compute calls here)."""
import ctypes
import os
import subprocess

import pytest

from starst3r_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from starst3r_b200 import build
        build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    protos = _lib.parse_header()
    assert len(protos) >= 10
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in include/starst3r_b200.h but not exported"


def test_no_torch_types_in_abi():
    import re
    text = re.sub(r"/\*.*?\*/", "", open(_lib.HEADER_PATH).read(), flags=re.S)   # prototypes only
    assert "torch" not in text.lower() and "at::" not in text and "Tensor" not in text and "std::" not in text


def test_dynamic_symbols_are_only_the_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    names = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert names and all(n.startswith("st3r_") for n in names), names


def test_version_and_sizes(lib):
    assert lib.st3r_abi_version() == 1
    # np.mgrid[S//2:H:S, S//2:W:S] seed counts (fast_nn.py:118-121)
    import numpy as np
    for H, W, S in [(512, 512, 8), (48, 64, 8), (50, 37, 8), (3, 3, 8), (4, 5, 8), (1072, 1920, 8), (17, 9, 4)]:
        ref = np.mgrid[S // 2:H:S, S // 2:W:S].reshape(2, -1).shape[1]
        assert lib.st3r_recip_seed_count(H, W, S) == ref
    assert lib.st3r_extract_corres_ws_bytes(512, 512, 512, 512, 8, 10) > 0
    assert lib.st3r_nn_argmax_ws_bytes(4096, 262144, 24) >= 4096 * 8


def test_bad_args_fail_loudly(lib):
    rc = lib.st3r_nn_argmax(None, 4, None, 4, 24, None, None, None, 0, 0, None)
    assert rc < 0 and b"null" in lib.st3r_last_error()
    rc = lib.st3r_nn_argmax(None, -1, None, 4, 24, None, None, None, 0, 0, None)
    assert rc < 0


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_cpu_device_is_refused():
    import torch
    from starst3r_b200 import match
    with pytest.raises(RuntimeError, match="CUDA"):
        match.fast_reciprocal_NNs(torch.zeros(8, 8, 24), torch.zeros(8, 8, 24), device="cpu", dist="dot")
