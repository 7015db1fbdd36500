"""ORACLE / TEST INFRASTRUCTURE — stand-in for `cormorant.nn` (risilab/cormorant @6a4b6370).

Restated layers used on the reference hot path (molgym/agents/covariant/modules.py:4-8,59-76,171-178):
InputLinear, NoLayer, RadialFilters/RadPolyTrig, CatMixReps (+ CatMixRepsScalar, MixReps, DotMatrix,
MaskLevel used by the edge/atom levels).  Upstream source is not available here: parity unpinned.

Named UNVERIFIED switches: #3 radial basis form / init, #4 DotMatrix & MaskLevel forms.
"""
import math

import torch
import torch.nn as nn

from . import so3_lib
from .cg_lib import CGModule
from .so3_lib import SO3Scalar, SO3Tau, SO3Vec, SO3Weight


class NoLayer(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()

    def forward(self, *args, **kwargs):
        return None

    @property
    def tau(self):
        return SO3Tau([])

    @property
    def num_scalars(self):
        return 0


class InputLinear(nn.Module):
    """Linear(num_in -> 2*num_out) read as num_out complex l=0 channels, zeroed on padded atoms."""

    def __init__(self, num_in, num_out, bias=True, device=None, dtype=torch.float):
        super().__init__()
        self.num_in = num_in
        self.num_out = num_out
        self.lin = nn.Linear(num_in, 2 * num_out, bias=bias)
        self.lin.to(device=device, dtype=dtype)
        self.zero = torch.tensor(0, dtype=dtype, device=device)

    def forward(self, atom_features, atom_mask, ignore, edge_mask, norms):
        atom_mask = atom_mask.unsqueeze(-1)
        out = torch.where(atom_mask, self.lin(atom_features), self.zero.to(atom_features.dtype))
        out = out.view(atom_features.shape[0:2] + (self.num_out, 1, 2))
        return SO3Vec([out])

    @property
    def tau(self):
        return SO3Tau([self.num_out])


class RadPolyTrig(nn.Module):
    """Radial basis sin(2 pi s r + phi) * r^-p with learnable (s, phi), mixed per ell by a Linear."""

    def __init__(self, max_sh, basis_set, num_channels, mix=False, device=None, dtype=torch.float):
        super().__init__()
        trig_basis, rpow = basis_set
        self.rpow = rpow
        self.max_sh = max_sh
        assert trig_basis >= 0 and rpow >= 0
        self.num_rad = (trig_basis + 1) * (rpow + 1)
        self.num_channels = num_channels

        scales = torch.cat([torch.arange(trig_basis + 1), torch.arange(trig_basis + 1)]).view(1, 1, 1, -1)
        phases = torch.cat([torch.zeros(trig_basis + 1), math.pi / 2 * torch.ones(trig_basis + 1)]).view(1, 1, 1, -1)
        scales = scales.to(device=device, dtype=dtype)
        phases = phases.to(device=device, dtype=dtype)
        phases[0, 0, 0, 0] = math.pi / 2  # avoid the dead sin(0*r + 0) feature
        self.scales = nn.Parameter(scales)
        self.phases = nn.Parameter(phases)

        self.mix = mix
        if mix == 'cplx' or mix is True:
            self.mix = 'cplx'
            self.linear = nn.ModuleList(
                [nn.Linear(2 * self.num_rad, 2 * self.num_channels) for _ in range(max_sh + 1)]).to(device=device,
                                                                                                    dtype=dtype)
            self.tau = SO3Tau((num_channels, ) * (max_sh + 1))
        elif mix == 'real':
            self.linear = nn.ModuleList(
                [nn.Linear(2 * self.num_rad, self.num_channels) for _ in range(max_sh + 1)]).to(device=device,
                                                                                                dtype=dtype)
            self.tau = SO3Tau((num_channels, ) * (max_sh + 1))
        elif mix == 'none' or mix is False:
            self.mix = 'none'
            self.linear = None
            self.tau = SO3Tau((self.num_rad, ) * (max_sh + 1))
        else:
            raise ValueError('Can only specify mix = real, cplx, or none! {}'.format(mix))
        self.zero = torch.tensor(0, device=device, dtype=dtype)

    def forward(self, norms, edge_mask):
        s = norms.shape
        zero = self.zero.to(norms.dtype)
        edge_mask = (edge_mask * (norms > 0)).unsqueeze(-1).bool()
        norms = norms.unsqueeze(-1)
        rad_powers = torch.stack([torch.where(edge_mask, norms.pow(-p), zero) for p in range(self.rpow + 1)], dim=-1)
        rad_trig = torch.where(edge_mask, torch.sin((2 * math.pi * self.scales) * norms + self.phases),
                               zero).unsqueeze(-1)
        rad_prod = (rad_powers * rad_trig).view(s + (1, 2 * self.num_rad))
        if self.mix == 'cplx':
            radial_functions = [linear(rad_prod).view(s + (self.num_channels, 2)) for linear in self.linear]
        elif self.mix == 'real':
            radial_functions = [linear(rad_prod).view(s + (self.num_channels, )) for linear in self.linear]
            radial_functions = [torch.stack([rad, torch.zeros_like(rad)], dim=-1) for rad in radial_functions]
        else:
            radial_functions = [rad_prod.view(s + (self.num_rad, 2))] * (self.max_sh + 1)
        return SO3Scalar(radial_functions)


class RadialFilters(nn.Module):
    def __init__(self, max_sh, basis_set, num_channels_out, num_levels, mix=True, device=None, dtype=torch.float):
        super().__init__()
        self.num_levels = num_levels
        self.max_sh = max_sh
        rad_funcs = [
            RadPolyTrig(max_sh[level], basis_set, num_channels_out[level], mix=mix, device=device, dtype=dtype)
            for level in range(self.num_levels)
        ]
        self.rad_funcs = nn.ModuleList(rad_funcs)
        self.tau = [rad_func.tau for rad_func in self.rad_funcs]
        self.num_rad_channels = self.tau[0][0]
        self.device = device
        self.dtype = dtype

    def forward(self, norms, base_mask):
        return [rad_func(norms, base_mask) for rad_func in self.rad_funcs]


class MixReps(CGModule):
    """Per-ell complex channel mixing with weights [tau_out, tau_in, 2]."""

    def __init__(self, tau_in, tau_out, real=False, weight_init='randn', gain=1, device=None, dtype=None):
        super().__init__(device=device, dtype=dtype)
        tau_in = SO3Tau(tau_in)
        if isinstance(tau_out, int):
            tau_out = SO3Tau([tau_out if t > 0 else 0 for t in tau_in])
        else:
            tau_out = SO3Tau(tau_out)
        self.tau_in = tau_in
        self.tau_out = tau_out
        self.real = real
        if weight_init == 'randn':
            weights = SO3Weight.randn(tau_in, tau_out, device=device, dtype=dtype)
            weights = [w for w in weights]
        elif weight_init == 'rand':
            weights = SO3Weight.rand(tau_in, tau_out, device=device, dtype=dtype)
            weights = [2 * w - 1 for w in weights]
        else:
            raise NotImplementedError('weight_init can only be randn or rand for now')
        weights = [(gain / max(w.shape)) * w for w in weights]
        self.weights = nn.ParameterList([nn.Parameter(w) for w in weights])

    @property
    def tau(self):
        return self.tau_out

    def forward(self, rep):
        if SO3Tau.from_rep(rep) != self.tau_in:
            raise ValueError('Tau of input rep does not match initialized tau! rep: {} tau: {}'.format(
                SO3Tau.from_rep(rep), self.tau_in))
        return so3_lib.mix(self.weights, rep)


class CatReps(nn.Module):
    def __init__(self, taus_in, maxl=None):
        super().__init__()
        self.taus_in = taus_in = [SO3Tau(tau) for tau in taus_in if tau]
        if maxl is None:
            maxl = max(tau.maxl for tau in taus_in)
        self.maxl = maxl
        self.tau_cat = SO3Tau(list(SO3Tau.cat(taus_in))[:maxl + 1])

    def forward(self, reps):
        reps = [rep for rep in reps if rep is not None]
        reps_taus_in = [rep.tau for rep in reps]
        if reps_taus_in != self.taus_in:
            raise ValueError('Tau of input reps does not match predefined version! got: {} expected: {}'.format(
                reps_taus_in, self.taus_in))
        reps = [rep.truncate(self.maxl) for rep in reps]
        return so3_lib.cat(reps)


class CatMixReps(CatReps):
    def __init__(self, taus_in, tau_out, maxl=None, real=False, weight_init='randn', gain=1, device=None,
                 dtype=None):
        super().__init__(taus_in, maxl=maxl)
        self.mix_reps = MixReps(self.tau_cat, tau_out, real=real, weight_init=weight_init, gain=gain, device=device,
                                dtype=dtype)
        self.tau = self.mix_reps.tau

    def forward(self, reps_in):
        return self.mix_reps(super().forward(reps_in))


CatMixRepsScalar = CatMixReps  # so3_lib.mix dispatches on the container type


class DotMatrix(CGModule):
    """Per channel and ell: sum_m (-1)^m a_i,m a_j,-m (complex); all ell concatenated and given to every ell."""

    def __init__(self, tau_in=None, cat=True, device=None, dtype=None):
        super().__init__(device=device, dtype=dtype)
        self.tau_in = SO3Tau(tau_in) if tau_in is not None else None
        self.cat = cat
        if self.tau_in is not None:
            if cat:
                self.tau = SO3Tau([sum(self.tau_in)] * len(self.tau_in))
            else:
                self.tau = SO3Tau(list(self.tau_in))
            self.signs = [
                torch.pow(-1, torch.arange(-ell, ell + 1).double()).to(device=device, dtype=dtype).unsqueeze(-1)
                for ell in range(len(self.tau_in) + 1)
            ]
            self.conj = torch.tensor([1., -1.]).to(device=device, dtype=dtype)
        else:
            self.tau = None
            self.signs = None

    def forward(self, reps):
        if self.tau_in is not None and self.tau_in != reps.tau:
            raise ValueError('Initialized tau not consistent with tau from forward! {} {}'.format(
                self.tau, reps.tau))
        reps1 = [part.unsqueeze(-4) for part in reps]
        reps2 = [part.unsqueeze(-5) for part in reps]
        reps2 = [part.flip(-2) * sign.to(part.dtype) for part, sign in zip(reps2, self.signs)]
        conj = self.conj.to(reps1[0].dtype)
        dot_r = [(p1 * p2 * conj).sum(dim=(-2, -1)) for p1, p2 in zip(reps1, reps2)]
        dot_i = [(p1 * p2.flip(-1)).sum(dim=(-2, -1)) for p1, p2 in zip(reps1, reps2)]
        dots = [torch.stack([r, i], dim=-1) for r, i in zip(dot_r, dot_i)]
        if self.cat:
            dots = torch.cat(dots, dim=-2)
            dots = [dots] * len(reps)
        return SO3Scalar(dots)


class MaskLevel(nn.Module):
    def __init__(self, num_channels, hard_cut_rad, soft_cut_rad, soft_cut_width, cutoff_type, gaussian_mask=False,
                 eps=1e-3, device=None, dtype=None):
        super().__init__()
        self.gaussian_mask = gaussian_mask
        self.num_channels = num_channels
        self.hard_cut_rad = None
        self.soft_cut_rad = None
        self.soft_cut_width = None
        if 'hard' in cutoff_type:
            self.hard_cut_rad = hard_cut_rad
        if any(t in cutoff_type for t in ('soft', 'soft_hard', 'learn', 'learn_rad', 'learn_width', 'learn_all')):
            self.soft_cut_rad = soft_cut_rad * torch.ones(num_channels, device=device, dtype=dtype).view((1, 1, 1, -1))
            self.soft_cut_width = soft_cut_width * torch.ones(num_channels, device=device,
                                                              dtype=dtype).view((1, 1, 1, -1))
        if ('learn_all' in cutoff_type) or ('learn_rad' in cutoff_type):
            self.soft_cut_rad = nn.Parameter(self.soft_cut_rad)
        if ('learn_all' in cutoff_type) or ('learn_width' in cutoff_type):
            self.soft_cut_width = nn.Parameter(self.soft_cut_width)
        self.dtype = dtype
        self.eps = torch.tensor(eps, device=device, dtype=dtype)

    def forward(self, edge_net, edge_mask, norms):
        if self.hard_cut_rad is not None:
            edge_mask = edge_mask * (norms < self.hard_cut_rad)
        edge_mask = edge_mask.to(norms.dtype).unsqueeze(-1)
        if self.soft_cut_rad is not None:
            cut_width = torch.max(self.eps, self.soft_cut_width.abs()).to(norms.dtype)
            cut_rad = torch.max(self.eps, self.soft_cut_rad.abs()).to(norms.dtype)
            if self.gaussian_mask:
                edge_mask = edge_mask * torch.exp(-(norms.unsqueeze(-1) / cut_rad).pow(2))
            else:
                edge_mask = edge_mask * torch.sigmoid((cut_rad - norms.unsqueeze(-1)) / cut_width)
        edge_mask = edge_mask.unsqueeze(-1)
        return edge_net * edge_mask
