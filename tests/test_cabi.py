"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/molgym_b200.h declares (no compute calls:
there is no GPU in the build container)."""
import ctypes
import os
import re

from molgym_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'molgym_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(mgb_[a-z_0-9]+)\s*\(', text)))


def test_header_and_python_bindings_agree():
    assert set(declared_symbols()) == set(_cabi.EXPORTS)


def test_cuda_library_exports_every_declared_symbol():
    path = build.build_cuda()
    lib = ctypes.CDLL(path)
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym
    lib.mgb_is_cuda_build.restype = ctypes.c_int
    assert lib.mgb_is_cuda_build() == 1
    lib.mgb_version.restype = ctypes.c_int
    assert lib.mgb_version() >= 100


def test_emulator_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(build.build_cusim())
    for sym in declared_symbols():
        assert hasattr(lib, sym), sym
    lib.mgb_is_cuda_build.restype = ctypes.c_int
    assert lib.mgb_is_cuda_build() == 0


def test_product_package_refuses_to_run_without_cuda():
    import pytest
    import torch

    from molgym_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    with pytest.raises(_lib.MissingCudaLibrary):
        _lib.require_cuda_device(None)
    with pytest.raises(_lib.MissingCudaLibrary):
        _lib.require_cuda_device('cpu')
