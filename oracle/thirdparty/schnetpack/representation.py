"""ORACLE / TEST INFRASTRUCTURE — restated `schnetpack.representation.SchNet` (0.3 defaults):
Embedding(100, F, padding_idx=0); GaussianSmearing(0, 5, 25); 3 x interaction
[filter MLP 25->128 (ssp)->128, x cosine cutoff; in2f F->128 (no bias); sum_j y_j W_ij; f2out 128->F (ssp);
dense F->F]; residual add.  Parity unpinned."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def shifted_softplus(x):
    return F.softplus(x) - math.log(2.0)


class Dense(nn.Linear):
    def __init__(self, in_features, out_features, bias=True, activation=None):
        self.activation = activation
        super().__init__(in_features, out_features, bias)

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, inputs):
        y = super().forward(inputs)
        if self.activation:
            y = self.activation(y)
        return y


class GaussianSmearing(nn.Module):
    def __init__(self, start=0.0, stop=5.0, n_gaussians=50, centered=False, trainable=False):
        super().__init__()
        offset = torch.linspace(start, stop, n_gaussians)
        widths = torch.FloatTensor((offset[1] - offset[0]) * torch.ones_like(offset))
        if trainable:
            self.width = nn.Parameter(widths)
            self.offsets = nn.Parameter(offset)
        else:
            self.register_buffer('width', widths)
            self.register_buffer('offsets', offset)
        self.centered = centered

    def forward(self, distances):
        coeff = -0.5 / torch.pow(self.width, 2)
        diff = distances[:, :, :, None] - self.offsets[None, None, None, :]
        return torch.exp(coeff * torch.pow(diff, 2))


class CosineCutoff(nn.Module):
    def __init__(self, cutoff=5.0):
        super().__init__()
        self.register_buffer('cutoff', torch.FloatTensor([cutoff]))

    def forward(self, distances):
        cutoffs = 0.5 * (torch.cos(distances * math.pi / self.cutoff) + 1.0)
        return cutoffs * (distances < self.cutoff).float()


def atom_distances(positions, neighbors, neighbor_mask=None):
    n_batch = positions.size()[0]
    idx_m = torch.arange(n_batch, device=positions.device, dtype=torch.long)[:, None, None]
    pos_xyz = positions[idx_m, neighbors[:, :, :], :]
    dist_vec = pos_xyz - positions[:, :, None, :]
    distances = torch.norm(dist_vec, 2, 3)
    if neighbor_mask is not None:
        tmp = torch.zeros_like(distances)
        tmp[neighbor_mask != 0] = distances[neighbor_mask != 0]
        distances = tmp
    return distances


class CFConv(nn.Module):
    def __init__(self, n_in, n_filters, n_out, filter_network, cutoff_network=None, activation=None):
        super().__init__()
        self.in2f = Dense(n_in, n_filters, bias=False, activation=None)
        self.f2out = Dense(n_filters, n_out, bias=True, activation=activation)
        self.filter_network = filter_network
        self.cutoff_network = cutoff_network

    def forward(self, x, r_ij, neighbors, pairwise_mask, f_ij=None):
        if f_ij is None:
            f_ij = r_ij.unsqueeze(-1)
        W = self.filter_network(f_ij)
        if self.cutoff_network is not None:
            C = self.cutoff_network(r_ij)
            W = W * C.unsqueeze(-1)
        y = self.in2f(x)
        nbh_size = neighbors.size()
        nbh = neighbors.reshape(-1, nbh_size[1] * nbh_size[2], 1)
        nbh = nbh.expand(-1, -1, y.size(2))
        y = torch.gather(y, 1, nbh)
        y = y.view(nbh_size[0], nbh_size[1], nbh_size[2], -1)
        y = y * W
        y = (y * pairwise_mask[..., None]).sum(dim=2)
        return self.f2out(y)


class SchNetInteraction(nn.Module):
    def __init__(self, n_atom_basis, n_spatial_basis, n_filters, cutoff):
        super().__init__()
        self.filter_network = nn.Sequential(
            Dense(n_spatial_basis, n_filters, activation=shifted_softplus),
            Dense(n_filters, n_filters),
        )
        self.cutoff_network = CosineCutoff(cutoff)
        self.cfconv = CFConv(n_atom_basis, n_filters, n_atom_basis, self.filter_network,
                             cutoff_network=self.cutoff_network, activation=shifted_softplus)
        self.dense = Dense(n_atom_basis, n_atom_basis, bias=True, activation=None)

    def forward(self, x, r_ij, neighbors, neighbor_mask, f_ij=None):
        v = self.cfconv(x, r_ij, neighbors, neighbor_mask, f_ij)
        return self.dense(v)


class SchNet(nn.Module):
    def __init__(self, n_atom_basis=128, n_filters=128, n_interactions=3, cutoff=5.0, n_gaussians=25, max_z=100):
        super().__init__()
        self.n_atom_basis = n_atom_basis
        self.embedding = nn.Embedding(max_z, n_atom_basis, padding_idx=0)
        self.distance_expansion = GaussianSmearing(0.0, cutoff, n_gaussians)
        self.interactions = nn.ModuleList([
            SchNetInteraction(n_atom_basis=n_atom_basis, n_spatial_basis=n_gaussians, n_filters=n_filters,
                              cutoff=cutoff) for _ in range(n_interactions)
        ])

    def forward(self, inputs):
        atomic_numbers = inputs['_atomic_numbers']
        positions = inputs['_positions']
        neighbors = inputs['_neighbors']
        neighbor_mask = inputs['_neighbor_mask']
        x = self.embedding(atomic_numbers)
        r_ij = atom_distances(positions, neighbors, neighbor_mask=neighbor_mask)
        f_ij = self.distance_expansion(r_ij)
        for interaction in self.interactions:
            v = interaction(x, r_ij, neighbors, neighbor_mask, f_ij=f_ij)
            x = x + v
        return x
