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


def build_molecules_loop(observations: List, actions: np.ndarray, zs: Sequence[int], canvas_size: int):
    """-> numbers[B,3,M] i32, positions[B,3,M,3] f32 (M = canvas_size + 1), bags[B,Z] f32.
    Molecule 0 = the canvas (agent.py:124-128); molecules 1/2 = canvas + the new atom for +/- dihedral (agent.py:163-177).
    One canvas at a time, exactly as the reference places atoms: the statement the batched build_molecules is tested against."""
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


_AUX1, _AUX0 = np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0])


def build_molecules(observations: List, actions: np.ndarray, zs: Sequence[int], canvas_size: int):
    """build_molecules_loop for the whole minibatch at once (float64 numpy over the batch axis: the per-canvas Python loop was
    3.6 ms of the 6.3 ms C1 step).  Same conventions: non-null atoms compacted in canvas order, focus indexes the compacted
    list, reference atoms = focus and its two nearest (stable order), auxiliary axes for fewer than three atoms."""
    B, M, Z = len(observations), canvas_size + 1, len(zs)
    numbers = np.zeros((B, 3, M), dtype=np.int32)
    positions = np.zeros((B, 3, M, 3), dtype=np.float32)
    bags = np.zeros((B, Z), dtype=np.float32)
    if B == 0:
        return numbers, positions, bags
    labels = np.array([[a[0] for a in canvas] for canvas, _ in observations], dtype=np.int64).reshape(B, -1)
    xyz = np.array([[a[1] for a in canvas] for canvas, _ in observations], dtype=np.float64).reshape(B, -1, 3)
    bags[:] = np.array([bag for _, bag in observations], dtype=np.float32).reshape(B, Z)
    if labels.size and (labels.min() < 0 or labels.max() >= Z):
        bad = labels[(labels < 0) | (labels >= Z)][0]
        raise RuntimeError(f'Invalid atomic number index: {bad}')
    N = labels.shape[1]
    z_of = np.asarray(zs, dtype=np.int64)[labels]                                  # [B, N] atomic numbers, 0 = empty slot
    valid = z_of != 0
    n = valid.sum(axis=1)
    order = np.argsort(~valid, axis=1, kind='stable')                              # non-null atoms first, canvas order kept
    rows = np.arange(B)[:, None]
    z_c = np.where(np.arange(N)[None, :] < n[:, None], z_of[rows, order], 0)
    p_c = np.where((np.arange(N)[None, :] < n[:, None])[:, :, None], xyz[rows, order], 0.0)
    numbers[:, :, :N] = z_c[:, None, :]
    positions[:, :, :N] = p_c[:, None, :, :]
    focus = np.rint(actions[:, 1].astype(np.float64)).astype(np.int64)
    element = np.rint(actions[:, 2].astype(np.float64)).astype(np.int64)
    dist, ang, dih = (actions[:, k].astype(np.float64) for k in (3, 4, 5))
    if np.any(focus > n):
        raise RuntimeError('Focus greater than number of atoms')
    if np.any((n > 0) & (focus >= n)):
        raise IndexError('focus index out of bounds for the atoms on the canvas')
    has = n > 0
    f = np.where(has, focus, 0)
    d2f = np.sqrt(np.sum(np.square(p_c - p_c[np.arange(B), f][:, None, :]), axis=2))
    d2f = np.where(np.arange(N)[None, :] < n[:, None], d2f, np.inf)
    near = np.argsort(d2f, axis=1, kind='stable')                                  # [B, N]
    idx = np.arange(B)
    p2 = p_c[idx, near[:, 0]]
    q1 = p_c[idx, near[:, min(1, N - 1)]]
    q2 = p_c[idx, near[:, min(2, N - 1)]]
    p1 = np.where((n == 1)[:, None], p2 + _AUX1, q1)
    p0 = np.where((n == 1)[:, None], p2 + _AUX0, np.where((n == 2)[:, None], p2 + q1 + _AUX0 + _AUX1, q2))
    safe = has[:, None]
    p1 = np.where(safe, p1, _AUX1)       # empty canvases: any non-degenerate frame (their new atom sits at the origin)
    p0 = np.where(safe, p0, _AUX0)
    p2 = np.where(safe, p2, 0.0)
    v_b = p2 - p1
    v_b = v_b / np.sqrt(np.sum(v_b * v_b, axis=1))[:, None]
    c_ab = np.cross(p1 - p0, v_b)
    c_ab = c_ab / np.sqrt(np.sum(c_ab * c_ab, axis=1))[:, None]
    x = (dist * np.cos(ang))[:, None]
    y = (dist * np.cos(dih) * np.sin(ang))[:, None]
    side = np.cross(c_ab, v_b)
    new_z = np.asarray(zs, dtype=np.int64)[element]
    for v, sign in ((1, 1.0), (2, -1.0)):
        zc = (dist * np.sin(sign * dih) * np.sin(ang))[:, None]
        new = np.where(safe, p2 - v_b * x + side * y + c_ab * zc, 0.0)
        numbers[idx, v, n] = new_z
        positions[idx, v, n] = new
    return numbers, positions, bags
