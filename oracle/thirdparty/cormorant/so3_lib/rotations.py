"""ORACLE / TEST INFRASTRUCTURE — stand-in for `cormorant.so3_lib.rotations`.

Wigner-D matrices from z-y-z Euler angles and the matching Cartesian rotation,
restated from the textbook formulae (Sakurai 3.8.33 for the small d).  The
convention is fixed so that for R = Rz(alpha) Ry(beta) Rz(gamma) the
*conjugated* unit-norm harmonics used for Cormorant's edge features obey
conjY_l(R r) = D_l . conjY_l(r) with D applied on the m index from the left —
which is what the reference's equivariance tests need
(tests/agents/covariant/test_agent.py:43-61, test_so3_tools.py:107-130).
"""
import math

import numpy as np
import torch


def _small_d(j, beta):
    d = np.zeros((2 * j + 1, 2 * j + 1))
    f = math.factorial
    c, s = math.cos(beta / 2), math.sin(beta / 2)
    for mp in range(-j, j + 1):
        for m in range(-j, j + 1):
            pref = math.sqrt(f(j + m) * f(j - m) * f(j + mp) * f(j - mp))
            tot = 0.0
            for k in range(max(0, m - mp), min(j + m, j - mp) + 1):
                den = f(j + m - k) * f(k) * f(mp - m + k) * f(j - mp - k)
                tot += (-1)**(mp - m + k) * c**(2 * j + m - mp - 2 * k) * s**(mp - m + 2 * k) / den
            d[mp + j, m + j] = pref * tot
    return d


def wigner_d_complex(j, alpha, beta, gamma):
    m = np.arange(-j, j + 1)
    d = _small_d(j, beta)
    big = np.exp(-1j * m[:, None] * alpha) * d * np.exp(-1j * m[None, :] * gamma)
    # Representation acting on conj(Y) under r -> R r (see module docstring; checked numerically
    # in tests/test_oracle_thirdparty.py::test_wigner_convention).
    return big


def wigner_d_list(maxl, alpha, beta, gamma, device=None, dtype=None):
    out = []
    for ell in range(maxl + 1):
        big = wigner_d_complex(ell, alpha, beta, gamma)
        out.append(torch.tensor(np.stack([big.real, big.imag], axis=-1), device=device, dtype=dtype))
    return out


def _rz(a):
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _ry(b):
    c, s = math.cos(b), math.sin(b)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def euler_rot(alpha, beta, gamma):
    return _rz(alpha) @ _ry(beta) @ _rz(gamma)


def gen_rot(maxl, angles=None, device=None, dtype=None):
    """Random rotation: (SO3WignerD, 3x3 rotation matrix tensor, angles)."""
    from . import SO3WignerD
    if angles is None:
        a, b, c = (torch.rand(3, dtype=torch.double) * 2 * math.pi).tolist()
        angles = (a, b / 2, c)
    alpha, beta, gamma = angles
    D = SO3WignerD(wigner_d_list(maxl, alpha, beta, gamma, device=device, dtype=dtype))
    R = torch.tensor(euler_rot(alpha, beta, gamma), device=device, dtype=dtype)
    return D, R, angles


def rotate_part(D, z, dir='left'):
    Dr, Di = D.unbind(-1)
    zr, zi = z.unbind(-1)
    if dir == 'left':
        mm = lambda d, x: torch.einsum('ij,...kj->...ki', d, x)  # noqa: E731
    else:
        mm = lambda d, x: torch.einsum('ji,...kj->...ki', d, x)  # noqa: E731
    return torch.stack([mm(Dr, zr) - mm(Di, zi), mm(Di, zr) + mm(Dr, zi)], dim=-1)


def rotate_rep(D_list, rep, dir='left'):
    out = []
    for part in rep:
        ell = (part.shape[-2] - 1) // 2
        out.append(rotate_part(D_list[ell].to(part.dtype), part, dir=dir))
    return out
