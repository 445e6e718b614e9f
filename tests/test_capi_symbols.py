"""The C-ABI library loads and exports every symbol include/aesfhe_b200.h declares
(no compute calls: there is no GPU in CI)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "aesfhe_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fhe_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from aes_fhe_b200 import _capi
    assert sorted(_capi.SIGNATURES) == declared_symbols()


def test_library_exports_every_symbol(cuda_lib):
    lib = ctypes.CDLL(str(cuda_lib))
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.fhe_launch_count is not None


def test_missing_library_is_loud(tmp_path):
    import pytest
    from aes_fhe_b200 import _capi
    with pytest.raises(_capi.FheError):
        _capi.load(tmp_path / "nope.so")


def test_engine_refuses_cpu_when_no_gpu():
    import pytest
    import torch
    from aes_fhe_b200.engine import Engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from aes_fhe_b200 import _capi
    with pytest.raises(_capi.FheError):
        Engine(max_level=4, log_coeff_count=None)
