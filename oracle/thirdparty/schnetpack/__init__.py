"""ORACLE / TEST INFRASTRUCTURE — stand-in for schnetpack==0.3 (requirements.txt:24).

Restates, from the published SchNet architecture (Schuett et al. 2018) and schnetpack 0.3's defaults, the two
names the reference uses (molgym/agents/internal/agent.py:6,37-38,128,177): `AtomsConverter` and
`representation.SchNet`.  Upstream source is not available here: parity unpinned (UNVERIFIED switch #5).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import representation  # noqa: F401,E402


class Properties:
    Z = '_atomic_numbers'
    R = '_positions'
    cell = '_cell'
    cell_offset = '_cell_offset'
    neighbors = '_neighbors'
    neighbor_mask = '_neighbor_mask'
    atom_mask = '_atom_mask'


class AtomsConverter:
    """ase.Atoms -> batch-of-one input dict; every other atom is a neighbour (SimpleEnvironmentProvider)."""

    def __init__(self, environment_provider=None, collect_triples=False, device=torch.device('cpu')):
        self.device = device

    def __call__(self, atoms):
        n = len(atoms)
        if n == 1:
            nbh = -np.ones((1, 1), dtype=np.int64)
        else:
            nbh = np.tile(np.arange(n, dtype=np.int64)[np.newaxis], (n, 1))
            nbh = nbh[~np.eye(n, dtype=bool)].reshape(n, n - 1)
        inputs = {
            Properties.Z: torch.tensor(np.asarray(atoms.numbers, dtype=np.int64)),
            Properties.R: torch.tensor(np.asarray(atoms.positions, dtype=np.float32)),
            Properties.cell: torch.zeros(3, 3, dtype=torch.float32),
            Properties.cell_offset: torch.zeros(n, nbh.shape[1], 3, dtype=torch.float32),
            Properties.neighbors: torch.tensor(nbh),
        }
        inputs[Properties.atom_mask] = torch.ones_like(inputs[Properties.Z]).float()
        mask = inputs[Properties.neighbors] >= 0
        inputs[Properties.neighbor_mask] = mask.float()
        inputs[Properties.neighbors] = inputs[Properties.neighbors] * inputs[Properties.neighbor_mask].long()
        return {key: value.unsqueeze(0).to(self.device) for key, value in inputs.items()}
