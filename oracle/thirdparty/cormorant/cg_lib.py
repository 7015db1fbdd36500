"""ORACLE / TEST INFRASTRUCTURE — stand-in for `cormorant.cg_lib` (risilab/cormorant @6a4b6370).

Restates the published Cormorant algorithm (Anderson, Hy, Kondor, NeurIPS 2019):
Clebsch-Gordan dictionary, channel-wise CG product (optionally aggregated over
the neighbour index), and complex spherical harmonics built by the
Y_l ~ CG(Y_{l-1} x Y_1) recursion.  The package itself is not in /root/reference
(requirements.txt:3) — parity unpinned except for the spherical-harmonic
known-answer vectors in tests/agents/covariant/test_sphs.py:18-55.

Named UNVERIFIED switches:
  REL_NORMALIZE_DEFAULT  — default `normalize` of SphericalHarmonicsRel (#1)
  path ordering in cg_product: l1 outer loop, l2 inner loop (#2)
"""
import math
from fractions import Fraction

import torch
import torch.nn as nn

from .so3_lib import SO3Tau, SO3Vec

REL_NORMALIZE_DEFAULT = False  # UNVERIFIED switch #1


# ----------------------------------------------------------------------------------------------
# Clebsch-Gordan coefficients (exact rational arithmetic under the square root, Racah's formula)
# ----------------------------------------------------------------------------------------------
def clebsch(j1, m1, j2, m2, j, m):
    if m1 + m2 != m or not (abs(j1 - j2) <= j <= j1 + j2):
        return 0.0
    if abs(m1) > j1 or abs(m2) > j2 or abs(m) > j:
        return 0.0
    f = math.factorial
    pref = Fraction((2 * j + 1) * f(j + j1 - j2) * f(j - j1 + j2) * f(j1 + j2 - j), f(j1 + j2 + j + 1))
    pref *= f(j + m) * f(j - m) * f(j1 - m1) * f(j1 + m1) * f(j2 - m2) * f(j2 + m2)
    tot = Fraction(0)
    for k in range(0, j1 + j2 - j + 1):
        args = (k, j1 + j2 - j - k, j1 - m1 - k, j2 + m2 - k, j - j2 + m1 + k, j - j1 - m2 + k)
        if min(args) < 0:
            continue
        den = 1
        for a in args:
            den *= f(a)
        tot += Fraction((-1)**k, den)
    return float(tot) * math.sqrt(pref)


def cg_matrix(l1, l2):
    """[(l1+l2+1)^2-(l1-l2)^2, (2l1+1)(2l2+1)] real matrix; rows grouped by l = |l1-l2| .. l1+l2, m = -l..l."""
    lmin, lmax = abs(l1 - l2), l1 + l2
    n1, n2 = 2 * l1 + 1, 2 * l2 + 1
    mat = torch.zeros((n1 * n2, n1 * n2), dtype=torch.double)
    for ell in range(lmin, lmax + 1):
        off = ell * ell - lmin * lmin
        for m1 in range(-l1, l1 + 1):
            for m2 in range(-l2, l2 + 1):
                m = m1 + m2
                if abs(m) <= ell:
                    mat[off + ell + m, (l1 + m1) * n2 + (l2 + m2)] = clebsch(l1, m1, l2, m2, ell, m)
    return mat


class CGDict:
    def __init__(self, maxl=None, transpose=True, device=None, dtype=torch.float):
        self.maxl = -1
        self.device = device
        self.dtype = dtype
        self._d = {}
        if maxl is not None:
            self.update_maxl(maxl)

    def update_maxl(self, maxl):
        for l1 in range(maxl + 1):
            for l2 in range(maxl + 1):
                if (l1, l2) not in self._d:
                    self._d[(l1, l2)] = cg_matrix(l1, l2).to(device=self.device, dtype=self.dtype)
        self.maxl = max(self.maxl, maxl)
        return self

    def to(self, device=None, dtype=None):
        self.device = device if device is not None else self.device
        self.dtype = dtype if dtype is not None else self.dtype
        self._d = {k: v.to(device=self.device, dtype=self.dtype) for k, v in self._d.items()}
        return self

    def keys(self):
        return self._d.keys()

    def __getitem__(self, key):
        return self._d[key]

    def __bool__(self):
        return self.maxl >= 0


class CGModule(nn.Module):
    """Base class holding (maxl, device, dtype, cg_dict)."""

    def __init__(self, cg_dict=None, maxl=None, device=None, dtype=None):
        super().__init__()
        self.device = device if device is not None else torch.device('cpu')
        self.dtype = dtype if dtype is not None else torch.float
        if cg_dict is None and maxl is not None:
            cg_dict = CGDict(maxl=maxl, device=self.device, dtype=self.dtype)
        elif cg_dict is not None and maxl is not None and cg_dict.maxl < maxl:
            cg_dict.update_maxl(maxl)
        self._cg_dict = cg_dict
        self._maxl = maxl if maxl is not None else (cg_dict.maxl if cg_dict is not None else None)

    @property
    def cg_dict(self):
        return self._cg_dict

    @property
    def maxl(self):
        return self._maxl


# ----------------------------------------------------------------------------------------------
# CG product
# ----------------------------------------------------------------------------------------------
def cg_product_tau(tau1, tau2, maxl=math.inf):
    tau1, tau2 = list(tau1), list(tau2)
    out = {}
    for l1, n1 in enumerate(tau1):
        for l2, n2 in enumerate(tau2):
            if n1 != n2:
                raise ValueError(f'CG product needs equal channel counts, got {n1} and {n2}')
            for ell in range(abs(l1 - l2), min(l1 + l2, maxl) + 1):
                out[ell] = out.get(ell, 0) + n1
    return SO3Tau([out.get(ell, 0) for ell in range(max(out.keys()) + 1)])


def complex_kron_product(z1, z2, aggregate=False):
    """[..., C, M1, 2] x [..., C, M2, 2] -> [..., C, M1*M2, 2]; aggregate sums over the neighbour (j) index."""
    if aggregate:
        b1, b2 = z1.shape[:-3], z2.shape[:-3]
        if len(b1) == 3 and len(b2) == 2:
            z2 = z2.unsqueeze(1)
        elif len(b1) == 2 and len(b2) == 3:
            z1 = z1.unsqueeze(1)
        else:
            raise ValueError(f'Batch size error! {b1} {b2}')
    z1r, z1i = z1.unsqueeze(-2).unbind(-1)  # [..., C, M1, 1]
    z2r, z2i = z2.unsqueeze(-3).unbind(-1)  # [..., C, 1, M2]
    zr = z1r * z2r - z1i * z2i
    zi = z1r * z2i + z1i * z2r
    z = torch.stack([zr, zi], dim=-1)
    z = z.reshape(z.shape[:-3] + (z.shape[-3] * z.shape[-2], 2))
    if aggregate:
        z = z.sum(dim=2)
    return z


def cg_product(cg_dict, rep1, rep2, maxl=math.inf, minl=0, aggregate=False, ignore_check=False):
    ells1 = [(p.shape[-2] - 1) // 2 for p in rep1]
    ells2 = [(p.shape[-2] - 1) // 2 for p in rep2]
    big_l = min(max(ells1) + max(ells2), maxl)
    new_rep = [[] for _ in range(big_l + 1)]
    for l1, part1 in zip(ells1, rep1):
        for l2, part2 in zip(ells2, rep2):
            lmin, lmax = max(abs(l1 - l2), minl), min(l1 + l2, big_l)
            if lmin > lmax:
                continue
            lo = lmin * lmin - (l1 - l2)**2
            hi = (lmax + 1)**2 - (l1 - l2)**2
            cg_mat = cg_dict[(l1, l2)][lo:hi, :]
            kron = complex_kron_product(part1, part2, aggregate=aggregate)  # [..., C, M1*M2, 2]
            dec = torch.matmul(cg_mat.to(kron.dtype), kron)  # [..., C, sum(2l+1), 2]
            pieces = dec.split([2 * ell + 1 for ell in range(lmin, lmax + 1)], dim=-2)
            for ell, piece in zip(range(lmin, lmax + 1), pieces):
                new_rep[ell].append(piece)
    return [torch.cat(parts, dim=-3) for parts in new_rep if len(parts) > 0]


class CGProduct(CGModule):
    def __init__(self, tau1=None, tau2=None, aggregate=False, minl=0, maxl=None, cg_dict=None, dtype=None,
                 device=None):
        super().__init__(cg_dict=cg_dict, maxl=maxl, device=device, dtype=dtype)
        self.tau1 = SO3Tau(tau1) if tau1 is not None else None
        self.tau2 = SO3Tau(tau2) if tau2 is not None else None
        self.aggregate = aggregate
        self.minl = minl

    @property
    def tau(self):
        return cg_product_tau(self.tau1, self.tau2, maxl=self.maxl)

    tau_out = tau

    def forward(self, rep1, rep2):
        return SO3Vec(cg_product(self.cg_dict, rep1, rep2, maxl=self.maxl, minl=self.minl,
                                 aggregate=self.aggregate))


# ----------------------------------------------------------------------------------------------
# Spherical harmonics
# ----------------------------------------------------------------------------------------------
def pos_to_rep(pos, conj=False):
    x, y, z = pos.unbind(-1)
    s = -1.0 if not conj else 1.0
    zero = torch.zeros_like(z)
    m_minus = torch.stack([x, s * y], -1) / math.sqrt(2.0)
    m_zero = torch.stack([z, zero], -1)
    m_plus = torch.stack([-x, s * y], -1) / math.sqrt(2.0)
    return torch.stack([m_minus, m_zero, m_plus], dim=-2).unsqueeze(-3)  # [..., 1, 3, 2]


def spherical_harmonics(cg_dict, pos, maxsh, normalize=True, conj=False, sh_norm='unit'):
    s = pos.shape[:-1]
    pos = pos.reshape(-1, 3)
    if normalize:
        norm = pos.norm(dim=-1, keepdim=True)
        pos = torch.where(norm > 0, pos / norm, torch.zeros_like(pos))
    psi0 = torch.zeros(pos.shape[0], 1, 1, 2, dtype=pos.dtype, device=pos.device)
    psi0[..., 0] = math.sqrt(1 / (4 * math.pi))
    harms = [psi0]
    if maxsh >= 1:
        psi1 = pos_to_rep(pos, conj=conj) * math.sqrt(3 / (4 * math.pi))
        harms.append(psi1)
    if maxsh >= 2:
        new_psi = psi1
        for ell in range(2, maxsh + 1):
            new_psi = cg_product(cg_dict, [new_psi], [psi1], minl=0, maxl=ell)[-1]
            # <l-1 0 1 0 | l 0>: row (l, m=0) of the (l-1, 1) block, column (m1=0, m2=0)
            row = ell * ell - (ell - 2)**2 + ell
            col = (ell - 1) * 3 + 1
            cg_coeff = cg_dict[(ell - 1, 1)][row, col].to(new_psi.dtype)  # stays a tensor, as upstream
            new_psi = new_psi * (math.sqrt((4 * math.pi * (2 * ell + 1)) / (3 * (2 * ell - 1))) / cg_coeff)
            harms.append(new_psi)
    harms = [part.reshape(s + part.shape[1:]) for part in harms]
    if sh_norm == 'qm':
        pass
    elif sh_norm == 'unit':
        harms = [part * math.sqrt((4 * math.pi) / (2 * ell + 1)) for ell, part in enumerate(harms)]
    else:
        raise ValueError(f'Incorrect choice of spherial harmonic normalization: {sh_norm}')
    return SO3Vec(harms)


def spherical_harmonics_rel(cg_dict, pos1, pos2, maxsh, normalize=True, conj=False, sh_norm='unit'):
    rel_pos = pos1.unsqueeze(-2) - pos2.unsqueeze(-3)
    rel_norms = rel_pos.norm(dim=-1, keepdim=True)
    harms = spherical_harmonics(cg_dict, rel_pos, maxsh, normalize=normalize, conj=conj, sh_norm=sh_norm)
    return harms, rel_norms.squeeze(-1)


class SphericalHarmonics(CGModule):
    def __init__(self, maxl, normalize=True, conj=False, sh_norm='unit', cg_dict=None, dtype=None, device=None):
        super().__init__(cg_dict=cg_dict, maxl=maxl, device=device, dtype=dtype)
        self.normalize = normalize
        self.sh_norm = sh_norm
        self.conj = conj

    def forward(self, pos):
        return spherical_harmonics(self.cg_dict, pos, self.maxl, self.normalize, self.conj, self.sh_norm)


class SphericalHarmonicsRel(CGModule):
    def __init__(self, maxl, normalize=None, conj=False, sh_norm='unit', cg_dict=None, dtype=None, device=None):
        super().__init__(cg_dict=cg_dict, maxl=maxl, device=device, dtype=dtype)
        self.normalize = REL_NORMALIZE_DEFAULT if normalize is None else normalize
        self.sh_norm = sh_norm
        self.conj = conj

    def forward(self, pos1, pos2):
        return spherical_harmonics_rel(self.cg_dict, pos1, pos2, self.maxl, self.normalize, self.conj,
                                       self.sh_norm)
