"""ORACLE / TEST INFRASTRUCTURE — stand-in for `cormorant.so3_lib` (risilab/cormorant @6a4b6370).

The upstream package is an un-vendored dependency of the reference
(requirements.txt:3) and is not installable here, so the containers below are
restated from its published design: an SO(3) "vector" is a python list over
ell of real tensors `[..., tau_ell, 2*ell+1, 2]` (trailing 2 = re/im), a
"scalar" is a list of `[..., tau_ell, 2]`, a "weight" a list of
`[tau_out, tau_in, 2]`.  Parity unpinned (no upstream source to diff against).

Surface used by the reference: molgym/agents/covariant/agent.py:7,82-83,281;
so3_tools.py:5; spherical_dists.py:8; tests/agents/covariant/test_agent.py:9,50,58.
"""
import itertools
import math
from typing import Iterable, List

import torch

from . import rotations  # noqa: F401  (re-exported: cormorant.so3_lib.rotations)


class SO3Tau:
    """Multiplicity list tau_ell (number of channels for each ell)."""

    def __init__(self, tau=()):
        if isinstance(tau, SO3Tau):
            tau = tau._tau
        self._tau = tuple(int(t) for t in tau)

    @property
    def maxl(self):
        return len(self._tau) - 1

    @property
    def channels(self):
        vals = set(self._tau)
        return next(iter(vals)) if len(vals) == 1 else None

    @staticmethod
    def from_rep(rep):
        if rep is None:
            return SO3Tau([])
        if isinstance(rep, SO3Scalar):
            return SO3Tau([p.shape[-2] for p in rep])
        return SO3Tau([p.shape[-3] for p in rep])

    @staticmethod
    def cat(taus):
        out = []
        for per_l in itertools.zip_longest(*[list(t) for t in taus], fillvalue=0):
            out.append(sum(per_l))
        return SO3Tau(out)

    def __iter__(self):
        return iter(self._tau)

    def __len__(self):
        return len(self._tau)

    def __getitem__(self, i):
        return self._tau[i]

    def __bool__(self):
        return len(self._tau) > 0

    def __eq__(self, other):
        return tuple(self) == tuple(SO3Tau(other)) if other is not None else False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self._tau)

    def __add__(self, other):
        return SO3Tau(list(self) + list(other))

    def __radd__(self, other):
        return SO3Tau(list(other) + list(self))

    def __repr__(self):
        return f'SO3Tau{list(self._tau)}'


class _SO3List:
    cdim = None  # channel dim

    def __init__(self, parts: Iterable[torch.Tensor]):
        self._parts: List[torch.Tensor] = list(parts)

    def __iter__(self):
        return iter(self._parts)

    def __len__(self):
        return len(self._parts)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self.__class__(self._parts[i])
        return self._parts[i]

    @property
    def maxl(self):
        return len(self._parts) - 1

    @property
    def ells(self):
        return list(range(len(self._parts)))

    @property
    def tau(self):
        return SO3Tau.from_rep(self)

    def truncate(self, maxl):
        return self.__class__(self._parts[:maxl + 1])

    @property
    def shapes(self):
        return [p.shape for p in self._parts]


def _cmul(ar, ai, br, bi):
    return torch.stack([ar * br - ai * bi, ar * bi + ai * br], dim=-1)


class SO3Vec(_SO3List):
    """list over ell of [..., tau, 2l+1, 2]"""
    cdim = -3

    def apply_wigner(self, wigner_d, dir='left'):
        return SO3Vec(rotations.rotate_rep(wigner_d, self, dir=dir))

    def __mul__(self, other):
        if isinstance(other, SO3Scalar):
            return other.__mul__(self)
        return SO3Vec([p * other for p in self._parts])

    __rmul__ = __mul__


class SO3Scalar(_SO3List):
    """list over ell of [..., tau, 2]"""
    cdim = -2

    def __mul__(self, other):
        if isinstance(other, SO3Vec):
            # complex scalar (per channel) times complex vector, channel-broadcast
            out = []
            for s, v in zip(self._parts, other):
                sr, si = s.unsqueeze(-2).unbind(-1)
                vr, vi = v.unbind(-1)
                out.append(_cmul(sr, si, vr, vi))
            return SO3Vec(out)
        if isinstance(other, SO3Scalar):
            out = []
            for a, b in zip(self._parts, other):
                ar, ai = a.unbind(-1)
                br, bi = b.unbind(-1)
                out.append(_cmul(ar, ai, br, bi))
            return SO3Scalar(out)
        return SO3Scalar([p * other for p in self._parts])

    __rmul__ = __mul__


class SO3Weight(_SO3List):
    """list over ell of [tau_out, tau_in, 2]"""

    @staticmethod
    def rand(tau_in, tau_out, device=None, dtype=None):
        return SO3Weight([torch.rand((t2, t1, 2), device=device, dtype=dtype) for t1, t2 in zip(tau_in, tau_out)])

    @staticmethod
    def randn(tau_in, tau_out, device=None, dtype=None):
        return SO3Weight([torch.randn((t2, t1, 2), device=device, dtype=dtype) for t1, t2 in zip(tau_in, tau_out)])


class SO3WignerD(_SO3List):
    """list over ell of complex [2l+1, 2l+1, 2] rotation matrices acting on the m index."""

    @staticmethod
    def euler(maxl, angles=None, device=None, dtype=None):
        if angles is None:
            a, b, c = (torch.rand(3, dtype=torch.double) * 2 * math.pi).tolist()
            angles = (a, b / 2, c)
        return SO3WignerD(rotations.wigner_d_list(maxl, *angles, device=device, dtype=dtype))


def cat(reps_list):
    """Concatenate along the channel dim, ell by ell; shorter reps simply stop contributing."""
    cls = reps_list[0].__class__
    per_l = [[p for p in parts if p is not None] for parts in itertools.zip_longest(*reps_list, fillvalue=None)]
    return cls([torch.cat(parts, dim=cls.cdim) for parts in per_l])


def mix(weights, rep):
    """Complex channel mixing W[t_out, t_in] applied ell by ell."""
    out = []
    if isinstance(rep, SO3Scalar):
        for w, p in zip(weights, rep):
            wr, wi = w.unbind(-1)
            pr, pi = p.unbind(-1)
            wr, wi = wr.t(), wi.t()
            out.append(torch.stack([pr @ wr - pi @ wi, pr @ wi + pi @ wr], dim=-1))
        return SO3Scalar(out)
    for w, p in zip(weights, rep):
        wr, wi = w.unbind(-1)
        pr, pi = p.unbind(-1)
        out.append(torch.stack([wr @ pr - wi @ pi, wi @ pr + wr @ pi], dim=-1))
    return SO3Vec(out)
