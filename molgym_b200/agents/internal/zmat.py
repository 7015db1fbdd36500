"""Z-matrix placement of a new atom (reference: molgym/agents/internal/zmat.py:66-133), float64 host arithmetic, and the
construction of the three SchNet "molecules" per observation that the C ABI's mgb_int_forward takes."""
from typing import List, Sequence

import numpy as np


def position_point(p0: np.ndarray, p1: np.ndarray, p2: np.ndarray, distance: float, angle: float, dihedral: float) -> np.ndarray:
    """zmat.py:66-96: the point `distance` from p2, at `angle` to p1 and `dihedral` to p0."""
    x = distance * np.cos(angle)
    y = distance * np.cos(dihedral) * np.sin(angle)
    z = distance * np.sin(dihedral) * np.sin(angle)
    v_b = p2 - p1
    v_b = v_b / np.linalg.norm(v_b)
    c_ab = np.cross(p1 - p0, v_b)
    c_ab = c_ab / np.linalg.norm(c_ab)
    return p2 - v_b * x + np.cross(c_ab, v_b) * y + c_ab * z


def position_atom_helper(positions: Sequence[np.ndarray], focus: int, distance: float, angle: float, dihedral: float) -> np.ndarray:
    """zmat.py:99-133: reference atoms = focus and its two nearest atoms (auxiliary axes when fewer than three exist)."""
    n = len(positions)
    if focus > n:
        raise RuntimeError('Focus greater than number of atoms')
    if n == 0:
        return np.zeros(3, dtype=np.float64)
    pos = np.asarray(positions, dtype=np.float64)
    order = np.argsort(np.sqrt(np.sum(np.square(pos - pos[focus]), axis=1)), kind='stable')
    aux1, aux0 = np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0])
    p2 = pos[order[0]]
    if n == 1:
        p1, p0 = p2 + aux1, p2 + aux0
    elif n == 2:
        p1 = pos[order[1]]
        p0 = p2 + p1 + aux0 + aux1
    else:
        p1, p0 = pos[order[1]], pos[order[2]]
    return position_point(p0, p1, p2, distance, angle, dihedral)


def build_molecules(observations: List, actions: np.ndarray, zs: Sequence[int], canvas_size: int):
    """-> numbers[B,3,M] i32, positions[B,3,M,3] f32 (M = canvas_size + 1), bags[B,Z] f32.
    Molecule 0 = the canvas (agent.py:124-128); molecules 1/2 = canvas + the new atom for +/- dihedral (agent.py:163-177)."""
    B, M = len(observations), canvas_size + 1
    numbers = np.zeros((B, 3, M), dtype=np.int32)
    positions = np.zeros((B, 3, M, 3), dtype=np.float32)
    bags = np.zeros((B, len(zs)), dtype=np.float32)
    for b, (canvas, bag) in enumerate(observations):
        pts, nums = [], []
        for label, xyz in canvas:
            if label < 0 or label >= len(zs):
                raise RuntimeError(f'Invalid atomic number index: {label}')
            if zs[label] != 0:
                nums.append(zs[label])
                pts.append(np.asarray(xyz, dtype=np.float64))
        n = len(nums)
        bags[b] = bag
        focus, element = int(round(float(actions[b, 1]))), int(round(float(actions[b, 2])))
        dist, ang, dih = float(actions[b, 3]), float(actions[b, 4]), float(actions[b, 5])
        for v, sign in enumerate((None, 1.0, -1.0)):
            numbers[b, v, :n] = nums
            if n:
                positions[b, v, :n] = np.asarray(pts, dtype=np.float64)
            if sign is not None:
                numbers[b, v, n] = zs[element]
                positions[b, v, n] = position_atom_helper(pts, focus, dist, ang, sign * dih)
    return numbers, positions, bags
