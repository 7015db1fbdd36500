"""ORACLE / TEST INFRASTRUCTURE — restated covariant actor-critic (reference: molgym/agents/covariant/*).

A functional restatement of `CovariantAC.step` in evaluate mode (actions given), built on the restated
cormorant stand-in.  Sub-module attribute names mirror the reference module tree so a reference
`state_dict()` loads directly.  Every block cites the reference lines it follows.  Works in float32
(the reference's dtype, agent.py:38) or float64 (the tolerance arbiter).
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
from cormorant.cg_lib import CGDict, CGProduct, SphericalHarmonics, SphericalHarmonicsRel
from cormorant.models.cormorant_cg import CormorantCG
from cormorant.nn import CatMixReps, InputLinear, NoLayer, RadialFilters
from cormorant.so3_lib import SO3Tau, SO3Vec
from scipy.integrate import lebedev_rule

ATOMIC_NUMBER_MAX = 103


def pack_observations(observations: Sequence, zs: Sequence[int], canvas_size: int):
    """ObservationType tuples -> (positions[B,N,3] f32, charges[B,N] i32, bags[B,Z] f32).
    Follows spaces.py:55-61 (null-symbol items are dropped, the rest compacted to the front),
    covariant/tools.py:8-49 (zero padding) and covariant/agent.py:191 (bags)."""
    B = len(observations)
    pos = np.zeros((B, canvas_size, 3), dtype=np.float32)
    charges = np.zeros((B, canvas_size), dtype=np.int32)
    bags = np.zeros((B, len(zs)), dtype=np.float32)
    for b, (canvas, bag) in enumerate(observations):
        k = 0
        for label, xyz in canvas:
            if label < 0:
                raise RuntimeError(f'Invalid atomic number: {label}')
            if zs[label] != 0:
                charges[b, k] = zs[label]
                pos[b, k] = np.asarray(xyz, dtype=np.float64).astype(np.float32)
                k += 1
        bags[b] = bag
    return pos, charges, bags


class MLP(nn.Module):
    """molgym/modules.py:30-50 — orthogonal weights, zero bias, ReLU between layers."""

    def __init__(self, input_dim, output_dims):
        super().__init__()
        dims = (input_dim, ) + tuple(output_dims)
        layers = []
        for a, b in zip(dims[:-1], dims[1:]):
            lin = nn.Linear(a, b)
            nn.init.orthogonal_(lin.weight.data)
            nn.init.constant_(lin.bias.data, 0)
            layers.append(lin)
        self.layers = nn.ModuleList(layers)

    def forward(self, x):
        for layer in self.layers[:-1]:
            x = torch.relu(layer(x))
        return self.layers[-1](x)


def masked_softmax(logits, mask):
    """molgym/modules.py:26-27 with torch_scatter.composite.scatter_softmax restated: softmax inside the
    mask==1 group (group max subtracted, eps=1e-12 on the sum), zero elsewhere."""
    neg = torch.finfo(logits.dtype).min
    mx = torch.where(mask, logits.detach(), torch.full_like(logits, neg)).max(dim=-1, keepdim=True)[0]
    e = torch.where(mask, (logits - mx).exp(), torch.zeros_like(logits))
    return e / (e.sum(dim=-1, keepdim=True) + 1e-12) * mask


def categorical_from_probs(probs):
    """torch.distributions.Categorical(probs=...) internals: renormalise, clamp to [eps, 1-eps], log."""
    probs = probs / probs.sum(-1, keepdim=True)
    eps = torch.finfo(probs.dtype).eps
    logits = torch.log(probs.clamp(min=eps, max=1 - eps))
    return probs, logits


def categorical_entropy(probs, logits):
    logits = torch.clamp(logits, min=torch.finfo(logits.dtype).min)
    return -(logits * probs).sum(-1)


def atomic_scalars(vec, maxl):
    """so3_tools.py:147-190 — [l=0 part (re, im)] + for each l [Re sum_m (-1)^m a_m a_-m, sum_m |a_m|^2]."""
    parts = [vec[0]]
    for ell, part in enumerate(vec):
        sign = torch.pow(-1.0, torch.arange(-ell, ell + 1, dtype=part.dtype))
        sign = torch.stack([sign, -sign], dim=-1)
        prod = (sign * part * part.flip(-2)).sum(dim=(-1, -2), keepdim=True)
        norm = (part * part).sum(dim=(-1, -2), keepdim=True)
        parts.append(torch.cat([prod, norm], dim=-1))
    return torch.cat(parts, dim=-3).flatten(start_dim=-3)


class OracleCormorant(nn.Module):
    """covariant/modules.py:11-135."""

    def __init__(self, maxl, num_cg_levels, num_channels, num_species, soft_cut_rad, charge_scale, bag_scale,
                 cg_dict, dtype):
        super().__init__()
        self.charge_power = 2  # agent.py:72
        self.charge_scale = charge_scale
        self.bag_scale = bag_scale
        self.num_species = num_species
        self.dtype = dtype
        lv = [maxl] * num_cg_levels
        self.sph_harms = SphericalHarmonicsRel(maxl=maxl, conj=True, dtype=dtype, cg_dict=cg_dict)  # :52-56
        self.rad_funcs = RadialFilters(max_sh=lv, basis_set=[3, 3], num_channels_out=num_channels,
                                       num_levels=num_cg_levels, dtype=dtype)  # :59-67
        num_scalars_in = num_species * (self.charge_power + 1) + num_species  # :69
        self.input_func_atom = InputLinear(num_scalars_in, num_channels[0], dtype=dtype)  # :72
        self.input_func_edge = NoLayer()
        self.cormorant_cg = CormorantCG(maxl=lv, max_sh=lv, tau_in_atom=self.input_func_atom.tau,
                                        tau_in_edge=self.input_func_edge.tau, tau_pos=self.rad_funcs.tau,
                                        num_cg_levels=num_cg_levels, num_channels=num_channels,
                                        level_gain=[10.0] * num_cg_levels, weight_init='rand',
                                        cutoff_type=['soft'], hard_cut_rad=[soft_cut_rad] * num_cg_levels,
                                        soft_cut_rad=[soft_cut_rad] * num_cg_levels,
                                        soft_cut_width=[0.2] * num_cg_levels, cat=True, gaussian_mask=False,
                                        dtype=dtype, cg_dict=cg_dict)  # :78-95

    def atom_scalars(self, charges, one_hot, bags):
        # prepare_input :116-135
        q = (charges.to(self.dtype).unsqueeze(-1) / self.charge_scale).pow(
            torch.arange(self.charge_power + 1, dtype=self.dtype))
        q = q.view(charges.shape + (1, self.charge_power + 1))
        q = (one_hot.to(self.dtype).unsqueeze(-1) * q).view(charges.shape[:2] + (-1, ))
        bag_tiled = (bags / self.bag_scale).unsqueeze(1).expand(q.shape[:-1] + (-1, ))
        return torch.cat([q, bag_tiled], dim=-1)

    def forward(self, positions, charges, one_hot, bags, atom_mask, edge_mask, all_levels=False):
        scalars = self.atom_scalars(charges, one_hot, bags)
        harmonics, norms = self.sph_harms(positions, positions)  # :102
        rad = self.rad_funcs(norms, edge_mask * (norms > 0))  # :103
        atoms_in = self.input_func_atom(scalars, atom_mask, None, edge_mask, norms)  # :106
        atoms_all, edges_all = self.cormorant_cg(atoms_in, atom_mask, None, edge_mask, rad, norms, harmonics)
        if all_levels:
            return atoms_all, edges_all, atoms_in
        return atoms_all[-1]  # :114


class OracleMixer(nn.Module):
    """covariant/modules.py:138-190 — (other x in) -> square -> concat [ag, sq, in] -> mix."""

    def __init__(self, channels, maxl, cg_dict, dtype):
        super().__init__()
        tau_in = SO3Tau([channels] * (maxl + 1))
        tau_other = SO3Tau([channels])
        self.cg_aggregate = CGProduct(tau_other, tau_in, maxl=maxl, dtype=dtype, cg_dict=cg_dict)
        tau_ag = list(self.cg_aggregate.tau)
        self.cg_power = CGProduct(tau_ag, tau_ag, maxl=maxl, dtype=dtype, cg_dict=cg_dict)
        tau_sq = list(self.cg_power.tau)
        self.cat_mix = CatMixReps([tau_ag, tau_sq, tau_in], channels, maxl=maxl, weight_init='rand', gain=10.0,
                                  dtype=dtype)

    def forward(self, atom_reps, other_reps):
        ag = self.cg_aggregate(other_reps, atom_reps)
        sq = self.cg_power(ag, ag)
        return self.cat_mix([ag, sq, atom_reps])


class CovariantOracle(nn.Module):
    """Restated CovariantAC (covariant/agent.py:20-334), evaluate mode."""

    def __init__(self, zs: List[int], canvas_size: int, min_max_distance: Tuple[float, float], network_width: int,
                 maxl: int, num_cg_levels: int, num_channels_hidden: int, num_channels_per_element: int,
                 num_gaussians: int, bag_scale: float, beta: Optional[float] = None, dtype=torch.float32):
        super().__init__()
        self.zs = list(zs)
        self.canvas_size = canvas_size
        self.dtype = dtype
        self.min_distance, self.max_distance = min_max_distance
        self.beta = beta
        self.max_sh = maxl
        self.cpe = num_channels_per_element
        self.num_gaussians = num_gaussians
        self.num_channels_out = len(zs) * num_channels_per_element  # agent.py:52
        cg_dict = CGDict(maxl=maxl, dtype=dtype)
        self.cg_model = OracleCormorant(maxl, num_cg_levels,
                                        [num_channels_hidden] * num_cg_levels + [self.num_channels_out], len(zs),
                                        min(self.max_distance, 2.1), max(zs), bag_scale, cg_dict, dtype)
        self.cg_mix = OracleMixer(num_channels_per_element, maxl, cg_dict, dtype)
        self.sph_harms = SphericalHarmonics(maxl=maxl, conj=False, sh_norm='qm', dtype=dtype, cg_dict=cg_dict)
        self.num_latent = (maxl + 2) * self.num_channels_out * 2  # so3_tools.py:167-171
        self.num_latent_element = (maxl + 2) * num_channels_per_element * 2
        self.phi_focus = MLP(self.num_latent, (network_width, 1))
        self.phi_element = MLP(self.num_latent, (network_width, len(zs)))
        self.phi_d = MLP(self.num_latent_element, (network_width, 2 * num_gaussians))
        self.distance_log_stds = nn.Parameter(torch.log(torch.tensor([0.1] * num_gaussians, dtype=dtype)))
        self.phi_trans = MLP(self.num_latent, (network_width, network_width))
        self.phi_v = MLP(network_width, (network_width, 1))
        self.to(dtype)
        pts, w = lebedev_rule(71)  # == quadpy lebedev_071 with weights / 4 pi (spherical_dists.py:209-212)
        self._leb_points = torch.tensor(pts.T.copy(), dtype=dtype)
        self._leb_logw = torch.log(torch.tensor(w / (4 * math.pi), dtype=dtype))

    # ------------------------------------------------------------------------------------------------
    def evaluate(self, positions, charges, bags, actions) -> Dict[str, torch.Tensor]:
        """positions [B,N,3], charges [B,N] int, bags [B,Z], actions [B,6] -> outputs + intermediates."""
        dt = self.dtype
        positions = torch.as_tensor(positions).to(dt)
        charges = torch.as_tensor(charges).to(torch.int32)
        bags = torch.as_tensor(bags).to(dt)
        actions = torch.as_tensor(actions).to(dt)
        B, N = charges.shape
        zs_t = torch.tensor(self.zs, dtype=dt)

        # parse_observations, agent.py:176-195
        one_hot = charges.unsqueeze(-1).to(dt) == zs_t.view(1, 1, -1)
        atom_mask = charges > 0
        edge_mask = atom_mask.unsqueeze(1) * atom_mask.unsqueeze(2)
        focus_mask = atom_mask.clone()
        focus_mask[:, 0] = True
        empty = ~atom_mask.any(dim=1)
        element_mask = bags > 0

        cov = self.cg_model(positions, charges, one_hot, bags, atom_mask, edge_mask)  # agent.py:217
        inv = atomic_scalars(cov, self.max_sh)  # agent.py:220

        # focus, agent.py:223-240
        focus_logits = self.phi_focus(inv).squeeze(-1)
        focus_p, focus_l = categorical_from_probs(masked_softmax(focus_logits, focus_mask))
        focus = torch.round(actions[:, 0]).long()
        rows = torch.arange(B)
        f_cov = [part[rows, focus] for part in cov]  # one-hot einsum == row pick
        f_inv = inv[rows, focus]

        # element, agent.py:243-259
        element_logits = self.phi_element(f_inv)
        element_p, element_l = categorical_from_probs(masked_softmax(element_logits, element_mask))
        element = torch.round(actions[:, 1]).long()
        idx = element.unsqueeze(-1) * self.cpe + torch.arange(self.cpe).unsqueeze(0)
        e_cov = SO3Vec([part[rows.unsqueeze(-1), idx] for part in f_cov])
        e_inv = atomic_scalars(e_cov, self.max_sh)

        # distance GMM, agent.py:263-276 + gmm.py:8-18 (MixtureSameFamily.log_prob)
        gmm_logits, mean_trans = self.phi_d(e_inv).split(self.num_gaussians, dim=-1)
        half_width = torch.tensor((self.max_distance - self.min_distance) / 2, dtype=dt)
        center = torch.tensor((self.min_distance + self.max_distance) / 2, dtype=dt)
        means = torch.tanh(mean_trans) * half_width + center
        stds = torch.exp(self.distance_log_stds).clamp(1e-6)
        d = actions[:, 2]
        comp = -((d.unsqueeze(-1) - means)**2) / (2 * stds**2) - stds.log() - math.log(math.sqrt(2 * math.pi))
        gmm_norm = gmm_logits - gmm_logits.logsumexp(dim=-1, keepdim=True)
        logp_d = torch.logsumexp(comp + torch.log_softmax(gmm_norm, dim=-1), dim=-1)

        # condition on distance, agent.py:279-282
        d_rep = torch.stack([d, torch.zeros_like(d)], dim=-1).view(B, 1, 1, 2).expand(B, self.cpe, 1, 2)
        cond = self.cg_mix(e_cov, SO3Vec([d_rep]))

        # spherical distribution, spherical_dists.py:79-286 + so3_tools.py:47-79
        k = sum((part.sum(dim=-3)**2).sum(dim=(-1, -2)) for part in cond)
        sqrt_k = k.clamp(min=1e-10).sqrt().view(B, 1, 1, 1)
        coeff = [part / sqrt_k for part in cond]
        orientation = actions[:, 3:6]

        def s_of(points):  # points [..., B, 3] -> sum_{l,tau,m} a Y  [..., B, 2]
            y = self.sph_harms(points)
            tot = 0
            for a, yl in zip(coeff, y):
                ar, ai = a.unbind(-1)
                yr, yi = yl.unbind(-1)
                tot = tot + torch.stack([ar * yr - ai * yi, ai * yr + ar * yi], dim=-1).sum(dim=(-3, -2))
            return tot

        if self.beta is not None:
            grid = self._leb_points.unsqueeze(-2)  # [G,1,3]
            log_unnorm_grid = -self.beta * (s_of(grid)**2).sum(-1)  # [G,B]
            log_z = math.log(4 * math.pi) + torch.logsumexp(log_unnorm_grid + self._leb_logw.unsqueeze(-1), dim=0)
            logp_o = -self.beta * (s_of(orientation)**2).sum(-1) - log_z
        else:
            log_z = torch.zeros(B, dtype=dt)
            p = (s_of(orientation)**2).sum(-1)
            p = torch.where(empty, torch.full_like(p, 1 / (4 * math.pi)), p)
            logp_o = torch.log(p.clamp(min=1e-10))

        logp_f = focus_l[rows, focus]
        logp_e = element_l[rows, element]
        logp = torch.stack([logp_f, logp_e, logp_d, logp_o], dim=-1).sum(-1)  # agent.py:295-301
        ent = categorical_entropy(focus_p, focus_l) + categorical_entropy(element_p, element_l)  # :304-308

        trans = self.phi_trans(inv)  # agent.py:313-316
        value_feats = torch.einsum('ba,baf->bf', atom_mask.to(dt), trans)
        v = self.phi_v(value_feats).squeeze(-1)

        return dict(logp=logp, ent=ent, v=v, logp_focus=logp_f, logp_element=logp_e, logp_distance=logp_d,
                    logp_orientation=logp_o, focus_probs=focus_p, element_probs=element_p, gmm_logits=gmm_norm,
                    gmm_means=means, gmm_stds=stds, log_z=log_z, coefficients=coeff, covariats=list(cov),
                    invariats=inv)

    def step(self, observations, actions):
        pos, charges, bags = pack_observations(observations, self.zs, self.canvas_size)
        return self.evaluate(pos, charges, bags, actions)
