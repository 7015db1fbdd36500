"""ORACLE / TEST INFRASTRUCTURE — stand-in for `cormorant.models.cormorant_cg.CormorantCG`
(called at molgym/agents/covariant/modules.py:78-95,110-111).  Restated from the Cormorant paper's
level structure: per level an edge network (dot-matrix + previous edge + radial, mixed and masked),
edge reps = edge scalars x Y, then an atom level (CG aggregate over neighbours, CG square, concat, mix).
Upstream source not available: parity unpinned."""
import torch.nn as nn

from ..cg_lib import CGModule, CGProduct
from ..nn import CatMixReps, CatMixRepsScalar, DotMatrix, MaskLevel


class CormorantEdgeLevel(nn.Module):
    def __init__(self, tau_atom, tau_edge, tau_pos, nout, max_sh, cutoff_type, hard_cut_rad, soft_cut_rad,
                 soft_cut_width, weight_init, gaussian_mask=False, device=None, dtype=None):
        super().__init__()
        self.dot_matrix = DotMatrix(tau_atom, cat=True, device=device, dtype=dtype)
        tau_dot = self.dot_matrix.tau
        edge_taus = [tau for tau in (tau_edge, tau_dot, tau_pos) if tau is not None]
        self.cat_mix = CatMixRepsScalar(edge_taus, nout, real=False, weight_init=weight_init, device=device,
                                        dtype=dtype)
        self.tau = self.cat_mix.tau
        self.mask_layer = MaskLevel(nout, hard_cut_rad, soft_cut_rad, soft_cut_width, cutoff_type,
                                    gaussian_mask=gaussian_mask, device=device, dtype=dtype)

    def forward(self, edge_in, atom_reps, pos_funcs, base_mask, norms):
        edge_dot = self.dot_matrix(atom_reps)
        edge_mix = self.cat_mix([edge_in, edge_dot, pos_funcs])
        return self.mask_layer(edge_mix, base_mask, norms)


class CormorantAtomLevel(CGModule):
    def __init__(self, tau_in, tau_pos, maxl, num_channels, level_gain, weight_init, device=None, dtype=None,
                 cg_dict=None):
        super().__init__(maxl=maxl, device=device, dtype=dtype, cg_dict=cg_dict)
        self.tau_in = tau_in
        self.tau_pos = tau_pos
        self.cg_aggregate = CGProduct(tau_pos, tau_in, maxl=self.maxl, aggregate=True, device=self.device,
                                      dtype=self.dtype, cg_dict=self.cg_dict)
        tau_ag = list(self.cg_aggregate.tau)
        self.cg_power = CGProduct(tau_in, tau_in, maxl=self.maxl, device=self.device, dtype=self.dtype,
                                  cg_dict=self.cg_dict)
        tau_sq = list(self.cg_power.tau)
        self.cat_mix = CatMixReps([tau_ag, tau_in, tau_sq], num_channels, maxl=self.maxl, weight_init=weight_init,
                                  gain=level_gain, device=self.device, dtype=self.dtype)
        self.tau = self.cat_mix.tau

    def forward(self, atom_reps, edge_reps, mask):
        reps_ag = self.cg_aggregate(edge_reps, atom_reps)
        reps_sq = self.cg_power(atom_reps, atom_reps)
        return self.cat_mix([reps_ag, atom_reps, reps_sq])


class CormorantCG(CGModule):
    def __init__(self, maxl, max_sh, tau_in_atom, tau_in_edge, tau_pos, num_cg_levels, num_channels, level_gain,
                 weight_init, cutoff_type, hard_cut_rad, soft_cut_rad, soft_cut_width, cat=True,
                 gaussian_mask=False, device=None, dtype=None, cg_dict=None):
        super().__init__(device=device, dtype=dtype, cg_dict=cg_dict)
        self.max_sh = max_sh
        atom_levels = nn.ModuleList()
        edge_levels = nn.ModuleList()
        tau_atom, tau_edge = tau_in_atom, tau_in_edge
        for level in range(num_cg_levels):
            edge_lvl = CormorantEdgeLevel(tau_atom, tau_edge, tau_pos[level], num_channels[level], max_sh[level],
                                          cutoff_type, hard_cut_rad[level], soft_cut_rad[level],
                                          soft_cut_width[level], weight_init, gaussian_mask=gaussian_mask,
                                          device=self.device, dtype=self.dtype)
            edge_levels.append(edge_lvl)
            tau_edge = edge_lvl.tau
            atom_lvl = CormorantAtomLevel(tau_atom, tau_edge, maxl[level], num_channels[level + 1],
                                          level_gain[level], weight_init, device=self.device, dtype=self.dtype,
                                          cg_dict=self.cg_dict)
            atom_levels.append(atom_lvl)
            tau_atom = atom_lvl.tau
        self.atom_levels = atom_levels
        self.edge_levels = edge_levels
        self.tau_levels_atom = [level.tau for level in atom_levels]
        self.tau_levels_edge = [level.tau for level in edge_levels]

    def forward(self, atom_reps, atom_mask, edge_net, edge_mask, rad_funcs, norms, sph_harm):
        assert len(self.atom_levels) == len(self.edge_levels) == len(rad_funcs)
        atoms_all, edges_all = [], []
        for idx, (atom_level, edge_level) in enumerate(zip(self.atom_levels, self.edge_levels)):
            edge_net = edge_level(edge_net, atom_reps, rad_funcs[idx], edge_mask, norms)
            edge_reps = edge_net * sph_harm
            atom_reps = atom_level(atom_reps, edge_reps, atom_mask)
            atoms_all.append(atom_reps)
            edges_all.append(edge_net)
        return atoms_all, edges_all
