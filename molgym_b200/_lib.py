"""Loader of the product C-ABI library (nvcc build for sm_100a).  There is no CPU fallback: a missing library, a
non-CUDA build or a machine without a CUDA device is an error."""
import ctypes
import os

from . import _cabi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libmolgym_b200.so')
_LIB = None


class MissingCudaLibrary(RuntimeError):
    pass


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise MissingCudaLibrary(f'{LIB_PATH} not found: build it with `python -m molgym_b200.build` '
                                 f'(or __graft_entry__.build()); molgym_b200 has no CPU fallback')
    lib = _cabi.bind(ctypes.CDLL(LIB_PATH))
    for sym in _cabi.EXPORTS:
        if not hasattr(lib, sym):
            raise MissingCudaLibrary(f'{LIB_PATH} does not export {sym}')
    if lib.mgb_is_cuda_build() != 1:
        raise MissingCudaLibrary(f'{LIB_PATH} is not the nvcc sm_100a build')
    _LIB = lib
    return lib


def require_cuda_device(device):
    import torch
    if not torch.cuda.is_available():
        raise MissingCudaLibrary('molgym_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    dev = torch.device(device if device is not None else 'cuda')
    if dev.type != 'cuda':
        raise MissingCudaLibrary(f'molgym_b200 agents run on CUDA devices only (got device={device!r})')
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device())
    return dev
