"""Minimal stand-ins for the parts of molgym/spaces.py the agents read (spaces.py:21-107): `observation_space.zs`,
`observation_space.canvas_space.size`, `action_space.zs`.  The reference's own gym-based ObservationSpace / ActionSpace
objects are accepted unchanged (duck typing); these exist so the package works without gym / ase installed."""
from typing import List, Sequence, Tuple

ObservationType = Tuple[tuple, tuple]


class CanvasItemSpace:
    def __init__(self, zs: Sequence[int]):
        self.zs = list(zs)


ActionSpace = CanvasItemSpace


class CanvasSpace:
    def __init__(self, size: int, zs: Sequence[int]):
        assert 0 in zs, '0 has to be in the list of atomic numbers'  # spaces.py:49
        self.size = size
        self.zs = list(zs)


class BagSpace:
    def __init__(self, zs: Sequence[int]):
        self.zs = list(zs)
        self.size = len(self.zs)


class ObservationSpace:
    def __init__(self, canvas_size: int, zs: List[int]):
        self.zs = list(zs)
        self.canvas_space = CanvasSpace(size=canvas_size, zs=zs)
        self.bag_space = BagSpace(zs=zs)
