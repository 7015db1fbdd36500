"""Observation packing: list of ObservationType tuples -> padded arrays (reference: CovariantAC.parse_observations,
agent.py:165-197; tools.process_atoms_list, covariant/tools.py:34-49; CanvasSpace.to_atoms, spaces.py:55-61).

The reference builds an ASE Atoms object per canvas and three device tensors per observation; here the nested tuples
are flattened once on the host and compacted by the C ABI's mgb_pack_observations."""
import ctypes
import os
from itertools import chain
from typing import Sequence

import numpy as np

from molgym_b200 import _cabi, _lib


_PACKER = None


def _native_packer():
    """The CPython extension built by molgym_b200.build.build_packer (host-side helper; optional)."""
    global _PACKER
    if _PACKER is None:
        import importlib.util
        from molgym_b200 import build
        path = build.packer_path()
        if os.path.exists(path):
            spec = importlib.util.spec_from_file_location('_mgb_packer', path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _PACKER = mod
        else:
            _PACKER = False
    return _PACKER


def flatten_observations(observations: Sequence, canvas_size: int, num_species: int):
    """-> labels[B,N] int32, xyz[B,N,3] float64, bags[B,Z] float32 exactly as stored in the tuples."""
    B = len(observations)
    native = _native_packer()
    if native:
        labels = np.empty((B, canvas_size), dtype=np.int32)
        xyz = np.empty((B, canvas_size, 3), dtype=np.float64)
        bags = np.empty((B, num_species), dtype=np.float32)
        native.flatten(observations, canvas_size, num_species, labels, xyz, bags)
        return labels, xyz, bags
    items = list(chain.from_iterable(obs[0] for obs in observations))
    if len(items) != B * canvas_size:
        raise RuntimeError(f'every canvas must hold exactly {canvas_size} items')
    labels = np.fromiter((it[0] for it in items), dtype=np.int32, count=B * canvas_size).reshape(B, canvas_size)
    xyz = np.fromiter(chain.from_iterable(it[1] for it in items), dtype=np.float64, count=B * canvas_size * 3)
    xyz = xyz.reshape(B, canvas_size, 3)
    bags = np.fromiter(chain.from_iterable(obs[1] for obs in observations), dtype=np.float32, count=B * num_species)
    return labels, xyz, bags.reshape(B, num_species)


def pack_observations(observations: Sequence, zs: Sequence[int], canvas_size: int, cfg=None, out=None, lib=None):
    """-> positions[B,N,3] f32, charges[B,N] i32 (null-symbol items dropped, real atoms compacted to the front, zero
    padding), bags[B,Z] f32.  `out` = (positions, charges, bags) arrays to fill (e.g. views of a pinned staging buffer)."""
    lib = lib if lib is not None else _lib.load()
    labels, xyz, bags = flatten_observations(observations, canvas_size, len(zs))
    B = len(observations)
    if out is not None:
        pos, charges, bags_out = out
        bags_out[...] = bags
        bags = bags_out
    else:
        pos = np.empty((B, canvas_size, 3), dtype=np.float32)
        charges = np.empty((B, canvas_size), dtype=np.int32)
    if cfg is None:
        cfg = _cabi.CovConfig()
        cfg.canvas_size, cfg.num_species = canvas_size, len(zs)
        for i, z in enumerate(zs):
            cfg.zs[i] = int(z)
    rc = lib.mgb_pack_observations(ctypes.byref(cfg), B, labels.ctypes.data, xyz.ctypes.data, pos.ctypes.data, charges.ctypes.data)
    if rc != 0:
        raise RuntimeError(f'Invalid observation: {lib.mgb_last_error().decode()}')
    return pos, charges, bags
