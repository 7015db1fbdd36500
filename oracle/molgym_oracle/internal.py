"""ORACLE / TEST INFRASTRUCTURE — restated internal-coordinate actor-critic (reference: molgym/agents/internal/agent.py,
molgym/agents/internal/zmat.py) in evaluate mode, on the restated schnetpack stand-in.  Attribute names mirror the
reference module tree so a reference `state_dict()` loads directly."""
import math
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
from schnetpack.representation import SchNet

from .covariant import MLP, categorical_entropy, categorical_from_probs, masked_softmax


def position_point(p0, p1, p2, distance, angle, dihedral):
    """zmat.py:66-96."""
    x = distance * np.cos(angle)
    y = distance * np.cos(dihedral) * np.sin(angle)
    z = distance * np.sin(dihedral) * np.sin(angle)
    v_a = p1 - p0
    v_b = p2 - p1
    v_b = v_b / np.linalg.norm(v_b)
    c_ab = np.cross(v_a, v_b)
    c_ab = c_ab / np.linalg.norm(c_ab)
    c_ab_b = np.cross(c_ab, v_b)
    return p2 - v_b * x + c_ab_b * y + c_ab * z


def position_atom_helper(positions: List[np.ndarray], focus: int, distance: float, angle: float, dihedral: float):
    """zmat.py:99-133."""
    if focus > len(positions):
        raise RuntimeError('Focus greater than number of atoms')
    if len(positions) == 0:
        return np.array([0, 0, 0], dtype=float)
    f = positions[focus]
    sorted_positions = sorted(positions, key=lambda p: np.sqrt(np.sum(np.square(p - f))))
    aux1, aux0 = np.array([1, 0, 0], dtype=float), np.array([0, 1, 0], dtype=float)
    if len(positions) == 1:
        p2 = sorted_positions[0]
        p1, p0 = p2 + aux1, p2 + aux0
    elif len(positions) == 2:
        p2, p1 = sorted_positions[0], sorted_positions[1]
        p0 = p2 + p1 + aux0 + aux1
    else:
        p2, p1, p0 = sorted_positions[0], sorted_positions[1], sorted_positions[2]
    return position_point(p0, p1, p2, distance=distance, angle=angle, dihedral=dihedral)


def schnet_inputs(numbers: Sequence[int], positions: np.ndarray):
    """schnetpack AtomsConverter with SimpleEnvironmentProvider (batch of one)."""
    n = len(numbers)
    if n == 1:
        nbh = -np.ones((1, 1), dtype=np.int64)
    else:
        nbh = np.tile(np.arange(n, dtype=np.int64)[np.newaxis], (n, 1))
        nbh = nbh[~np.eye(n, dtype=bool)].reshape(n, n - 1)
    nbh_t = torch.tensor(nbh)
    mask = (nbh_t >= 0).float()
    return {'_atomic_numbers': torch.tensor(np.asarray(numbers, dtype=np.int64)).unsqueeze(0),
            '_positions': torch.tensor(np.asarray(positions, dtype=np.float32)).unsqueeze(0),
            '_neighbors': (nbh_t * mask.long()).unsqueeze(0), '_neighbor_mask': mask.unsqueeze(0)}


class MLP3(MLP):
    pass


class SchNetOracle(nn.Module):
    """Restated SchNetAC (internal/agent.py:17-353), evaluate mode (actions given)."""

    def __init__(self, zs: List[int], canvas_size: int, min_max_distance: Tuple[float, float], network_width: int):
        super().__init__()
        self.zs = list(zs)
        self.num_atoms = canvas_size
        self.num_zs = len(zs)
        self.num_afeats = network_width // 2
        self.num_latent_beta = network_width // 4
        self.num_latent = self.num_afeats + self.num_latent_beta
        self.embedding_fn = SchNet(n_atom_basis=self.num_afeats)
        self.phi_beta = MLP(self.num_zs, (network_width, self.num_latent_beta))
        self.phi_focus = MLP(self.num_latent, (network_width, 1))
        self.phi_element = MLP(self.num_latent, (network_width, self.num_zs))
        self.phi_continuous = MLP(self.num_latent + self.num_zs, (network_width, 3))
        self.phi_kappa = MLP(self.num_latent, (network_width, 1))
        self.log_stds = nn.Parameter(torch.log(torch.tensor([0.15, 0.25, 0.25], dtype=torch.float32)))
        self.min_distance, self.max_distance = min_max_distance
        self.action_width = torch.tensor([self.max_distance - self.min_distance, math.pi, math.pi])
        self.action_center = 0.5 * torch.tensor([self.max_distance + self.min_distance, math.pi, math.pi])
        self.critic = MLP(self.num_latent, (network_width, network_width, 1))

    def _parse(self, observation):
        canvas, bag = observation
        numbers, positions = [], []
        for label, xyz in canvas:
            if self.zs[label] != 0:
                numbers.append(self.zs[label])
                positions.append(np.asarray(xyz, dtype=float))
        return numbers, positions, list(bag)

    def _embed(self, numbers, positions):
        return self.embedding_fn(schnet_inputs(numbers, np.asarray(positions)))[0]

    def step(self, observations, actions):
        B, N = len(observations), self.num_atoms
        actions_t = torch.as_tensor(np.asarray(actions))
        feats = torch.zeros(B, N, self.num_afeats)
        focus_mask = torch.zeros(B, N, dtype=torch.bool)
        counts = torch.zeros(B, self.num_zs)
        action_mask = torch.zeros(B, 6)
        parsed = [self._parse(o) for o in observations]
        rows = []
        for i, (numbers, positions, bag) in enumerate(parsed):
            n = len(numbers)
            if n > 0:
                rows.append((i, self._embed(numbers, positions)))
                focus_mask[i, :n] = True
            else:
                focus_mask[i, :1] = True
            counts[i] = torch.tensor(bag, dtype=torch.float32)
            action_mask[i] = torch.tensor([n >= 1, 1.0, n >= 1, n >= 2, n >= 3, n >= 3], dtype=torch.float32)
        for i, f in rows:   # agent.py:126-128
            feats = feats.clone()
            feats[i, :f.shape[0]] = f
        element_mask = counts > 0
        latent_bag = self.phi_beta(counts)
        latent = torch.cat([feats, latent_bag.unsqueeze(1).expand(-1, N, -1)], dim=-1)
        focus_logits = self.phi_focus(latent).squeeze(-1)
        focus_p, focus_l = categorical_from_probs(masked_softmax(focus_logits, focus_mask))
        focus = torch.round(actions_t[:, 1]).long()
        ar = torch.arange(B)
        focused = latent[ar, focus]
        element_p, element_l = categorical_from_probs(masked_softmax(self.phi_element(focused), element_mask))
        element = torch.round(actions_t[:, 2]).long()
        element_oh = torch.nn.functional.one_hot(element, self.num_zs).float()
        means = torch.tanh(self.phi_continuous(torch.cat([focused, element_oh], dim=-1))) * self.action_width / 2 + self.action_center
        scales = torch.exp(1e-6 + self.log_stds)
        cont = actions_t[:, 3:6].float()
        logp_cont = -((cont - means)**2) / (2 * scales**2) - scales.log() - math.log(math.sqrt(2 * math.pi))
        latent_bag_next = self.phi_beta(counts - element_oh)
        nxt = []
        for sign in (1.0, -1.0):
            fs = []
            for i, (numbers, positions, bag) in enumerate(parsed):
                new_pos = position_atom_helper(positions, int(round(float(actions_t[i, 1]))), float(actions_t[i, 3]),
                                               float(actions_t[i, 4]), sign * float(actions_t[i, 5]))
                z_new = self.zs[int(round(float(actions_t[i, 2])))]
                fs.append(self._embed(numbers + [z_new], positions + [new_pos])[-1])
            nxt.append(torch.stack(fs))
        v0 = self.phi_kappa(torch.cat([nxt[0], latent_bag_next], dim=-1))
        v1 = self.phi_kappa(torch.cat([nxt[1], latent_bag_next], dim=-1))
        kappa_logits = torch.cat([v0, v1], dim=-1)
        kappa_l = kappa_logits - kappa_logits.logsumexp(dim=-1, keepdim=True)
        kappa = torch.round(actions_t[:, 6]).long()
        sum_feats = (focus_mask.float().unsqueeze(-1) * feats).sum(dim=1)
        v = self.critic(torch.cat([sum_feats, latent_bag], dim=-1)).squeeze(-1)
        logp_terms = torch.stack([focus_l[ar, focus], element_l[ar, element], logp_cont[:, 0], logp_cont[:, 1], logp_cont[:, 2],
                                  kappa_l[ar, kappa]], dim=-1) * action_mask
        ent = categorical_entropy(focus_p, focus_l) * action_mask[:, 0] + categorical_entropy(element_p, element_l) * action_mask[:, 1]
        return dict(logp=logp_terms.sum(-1), ent=ent, v=v, logp_terms=logp_terms, focus_probs=focus_p, element_probs=element_p,
                    means=means, kappa_logits=kappa_logits, feats=feats, feats_next=nxt)
