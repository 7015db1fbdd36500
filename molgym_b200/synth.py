"""Synthetic PPO-buffer canvases for the configs named in BASELINE.json.

Host-side numpy only.  Observations are built exactly as the reference's spaces do
(molgym/spaces.py:63-74,103-104): `canvas` = tuple of `canvas_size` items `(label_index, (x, y, z))`, real atoms
first, padding `(zs.index(0), (0, 0, 0))`; `bag` = tuple of counts aligned with `zs`.  Geometry follows the
environment's placement rule (new atom at a sampled distance from an existing atom, rejected while any pair is
closer than 0.6 A — molgym/environment.py:27,91-98).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass
class WorkloadConfig:
    name: str
    model: str  # 'covariant' | 'internal'
    zs: List[int]
    canvas_size: int
    mini_batch_size: int
    bag: Dict[int, int]  # atomic number -> count (the formula multiset)
    min_max_distance: Tuple[float, float] = (0.8, 1.8)  # arg_parser.py:53-54
    bag_scale: int = 5
    beta: Optional[float] = None
    network_width: int = 128  # arg_parser.py:55-61
    maxl: int = 4
    num_cg_levels: int = 3
    num_channels_hidden: int = 10
    num_channels_per_element: int = 4
    num_gaussians: int = 3
    seed: int = 1000
    extra: dict = field(default_factory=dict)

    def agent_kwargs(self):
        if self.model == 'internal':
            return dict(min_max_distance=self.min_max_distance, network_width=self.network_width)
        return dict(min_max_distance=self.min_max_distance, network_width=self.network_width, maxl=self.maxl,
                    num_cg_levels=self.num_cg_levels, num_channels_hidden=self.num_channels_hidden,
                    num_channels_per_element=self.num_channels_per_element, num_gaussians=self.num_gaussians,
                    bag_scale=self.bag_scale, beta=self.beta)


# BASELINE.json `configs`, index-aligned (C1..C5 in SURVEY.md section 8).
CONFIGS = {
    'C1': WorkloadConfig('C1-SF6-internal', 'internal', [0, 9, 16], 7, 28, {16: 1, 9: 6}, (1.10, 2.10), 5, None,
                         seed=1001),
    'C2': WorkloadConfig('C2-SF6-covariant', 'covariant', [0, 9, 16], 7, 140, {16: 1, 9: 6}, (1.10, 2.10), 5, -10.0,
                         seed=1002),
    'C3': WorkloadConfig('C3-C3H5NO3-covariant', 'covariant', [0, 1, 6, 7, 8], 12, 1024, {6: 3, 1: 5, 7: 1, 8: 3},
                         (0.8, 1.8), 12, -10.0, seed=1003),
    'C4': WorkloadConfig('C4-stochastic-CHNO-covariant', 'covariant', [0, 1, 6, 7, 8], 22, 4096,
                         {6: 7, 1: 10, 7: 2, 8: 3}, (0.8, 1.8), 22, -10.0, seed=1004),
    'C5': WorkloadConfig('C5-solvation-covariant', 'covariant', [0, 1, 6, 8], 40, 8192, {6: 6, 1: 24, 8: 10},
                         (0.8, 1.8), 40, -10.0, seed=1005),
}


def _grow_geometry(rng, n, dmin, dmax, min_sep=0.6):
    pos = np.zeros((n, 3), dtype=np.float64)
    for k in range(1, n):
        for _ in range(200):
            anchor = pos[rng.integers(0, k)]
            direction = rng.normal(size=3)
            direction /= np.linalg.norm(direction)
            cand = anchor + rng.uniform(dmin, dmax) * direction
            if np.all(np.linalg.norm(pos[:k] - cand, axis=1) >= min_sep):
                break
        pos[k] = cand
    return pos


def make_observations(cfg: WorkloadConfig, batch: Optional[int] = None, seed: Optional[int] = None,
                      start_index: int = 0):
    """-> (observations, n_atoms[B]).  Canvas b holds (start_index + b) mod K atoms, K = min(N, bag size):
    occupancies 0..K-1 uniformly, empty canvas included, as complete episodes leave them in a PPO buffer."""
    rng = np.random.default_rng(cfg.seed if seed is None else seed)
    batch = cfg.mini_batch_size if batch is None else batch
    bag_atoms = [z for z, c in cfg.bag.items() for _ in range(c)]
    K = min(cfg.canvas_size, len(bag_atoms))
    null_index = cfg.zs.index(0)
    dmin, dmax = cfg.min_max_distance
    observations, n_atoms = [], []
    for b in range(batch):
        n = (start_index + b) % K
        order = rng.permutation(len(bag_atoms))
        placed = [bag_atoms[i] for i in order[:n]]
        pos = _grow_geometry(rng, n, dmin, dmax)
        remaining: Dict[int, int] = {z: 0 for z in cfg.zs}
        for i in order[n:]:
            remaining[bag_atoms[i]] += 1
        canvas = tuple((cfg.zs.index(z), tuple(float(x) for x in p)) for z, p in zip(placed, pos))
        canvas = canvas + ((null_index, (0.0, 0.0, 0.0)), ) * (cfg.canvas_size - n)
        bag = tuple(int(remaining[z]) for z in cfg.zs)
        observations.append((canvas, bag))
        n_atoms.append(n)
    return observations, np.asarray(n_atoms, dtype=np.int32)


def make_actions(cfg: WorkloadConfig, observations, n_atoms, seed: Optional[int] = None) -> np.ndarray:
    """Valid evaluate-mode actions.  Covariant: [focus, element, distance, ox, oy, oz] (covariant/agent.py:230-288);
    internal: [stop, focus, element, distance, angle, dihedral, kappa] (internal/agent.py:215-308)."""
    rng = np.random.default_rng((cfg.seed if seed is None else seed) + 7919)
    B = len(observations)
    dmin, dmax = cfg.min_max_distance
    focus = np.array([rng.integers(0, max(n, 1)) for n in n_atoms], dtype=np.float32)
    element = np.zeros(B, dtype=np.float32)
    for b, (_, bag) in enumerate(observations):
        avail = [i for i, c in enumerate(bag) if c > 0]
        element[b] = rng.choice(avail)
    distance = rng.uniform(dmin, dmax, size=B).astype(np.float32)
    if cfg.model == 'internal':
        angle = rng.uniform(0.3, np.pi - 0.3, size=B).astype(np.float32)
        dihedral = rng.uniform(0.1, np.pi - 0.1, size=B).astype(np.float32)
        kappa = rng.integers(0, 2, size=B).astype(np.float32)
        return np.stack([np.zeros(B, np.float32), focus, element, distance, angle, dihedral, kappa], axis=-1)
    o = rng.normal(size=(B, 3))
    o = (o / np.linalg.norm(o, axis=-1, keepdims=True)).astype(np.float32)
    return np.concatenate([np.stack([focus, element, distance], axis=-1), o], axis=-1).astype(np.float32)


def make_ppo_targets(cfg: WorkloadConfig, logp0: np.ndarray, seed: Optional[int] = None):
    """adv ~ N(0,1) standardised and ret ~ N(0,0.3) as float64 (buffer.py:106-114 hands ppo float64),
    old_logp = logp0 + N(0, 0.05) as float32 (buffer.py:116)."""
    rng = np.random.default_rng((cfg.seed if seed is None else seed) + 104729)
    B = len(logp0)
    adv = rng.normal(size=B)
    adv = (adv - adv.mean()) / (adv.std() + 1e-12)
    ret = rng.normal(scale=0.3, size=B)
    old_logp = (np.asarray(logp0, dtype=np.float64) + rng.normal(scale=0.05, size=B)).astype(np.float32)
    return old_logp, adv.astype(np.float64), ret.astype(np.float64)
