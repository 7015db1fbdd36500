"""Rollout mode (actions=None) and the `dists` objects of the response dict.

Reference: agent.py:229-292 (sample when self.training else argmax), spherical_dists.py:44-286 (rejection samplers on
S^2), gmm.py:8-27, so3_tools.py:8-58.  The sub-actions of a rollout step are drawn on the device by one kernel
(`rollout` below -> mgb_cov_rollout).  The distribution objects handed out in `response['dists']` keep the reference's
surface (`probs`, `coefficients`, `log_prob`, `sample`, `argmax`) as plain torch code for callers that inspect them; the
hot path does not run through them.
"""
import math

import numpy as np
import torch
import torch.distributions as D

_FACT = [1.0, 1.0, 2.0, 6.0, 24.0, 120.0, 720.0, 5040.0, 40320.0]


def sph_harm_qm(xyz: torch.Tensor) -> torch.Tensor:
    """Complex Y_lm (l <= 4, Condon-Shortley, orthonormal 'qm' norm) of the normalised vectors xyz [..., 3] ->
    [..., 25, 2] with lm = l*l + l + m (same closed forms as csrc/common.cuh::sph_harm_l4)."""
    nrm = xyz.norm(dim=-1, keepdim=True)
    v = torch.where(nrm > 0, xyz / nrm, torch.zeros_like(xyz))
    x, y, z = v.unbind(-1)
    r2 = x * x + y * y + z * z
    z2 = z * z
    one = torch.ones_like(z)
    er = [one, x, x * x - y * y]
    ei = [torch.zeros_like(z), y, 2 * x * y]
    er.append(er[2] * x - ei[2] * y); ei.append(er[2] * y + ei[2] * x)
    er.append(er[2] * er[2] - ei[2] * ei[2]); ei.append(2 * er[2] * ei[2])
    d = {(0, 0): one, (1, 0): z, (1, 1): one, (2, 0): 0.5 * (3 * z2 - r2), (2, 1): 3 * z, (2, 2): 3 * one,
         (3, 0): 0.5 * z * (5 * z2 - 3 * r2), (3, 1): 0.5 * (15 * z2 - 3 * r2), (3, 2): 15 * z, (3, 3): 15 * one,
         (4, 0): 0.125 * (35 * z2 * z2 - 30 * z2 * r2 + 3 * r2 * r2), (4, 1): 0.5 * z * (35 * z2 - 15 * r2),
         (4, 2): 0.5 * (105 * z2 - 15 * r2), (4, 3): 105 * z, (4, 4): 105 * one}
    out = [None] * 25
    for l in range(5):
        for m in range(l + 1):
            a = math.sqrt((2 * l + 1) / (4 * math.pi) * _FACT[l - m] / _FACT[l + m]) * d[(l, m)]
            sg = -1.0 if m % 2 else 1.0
            out[l * l + l + m] = torch.stack([sg * a * er[m], sg * a * ei[m]], dim=-1)
            if m > 0:
                out[l * l + l - m] = torch.stack([a * er[m], -a * ei[m]], dim=-1)
    return torch.stack(out, dim=-2)


def fibonacci_grid(n: int) -> np.ndarray:
    """so3_tools.py:8-20."""
    golden = (1 + 5**0.5) / 2
    index = np.arange(0, n)
    theta = np.arccos(1 - 2 * (index + 0.5) / n)
    phi = 2 * np.pi * index / golden
    return np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], axis=-1)


class SO3DistributionView:
    """The fourth entry of `dists`: exposes `.coefficients` (list over l of [B, tau, 2l+1, 2]) and `log_prob`, `prob`,
    `sample`, `argmax` with the reference's semantics (SO3Distribution when beta is None, else ExpSO3Distribution)."""

    def __init__(self, coeff: torch.Tensor, beta, log_z: torch.Tensor, empty: torch.Tensor):
        # coeff: [B, 25, tau, 2] normalised a_lm
        self._coeff = coeff
        self.beta = beta
        self.log_z = log_z if beta is not None else None
        self.empty = empty
        self.batch_shape = torch.Size((coeff.shape[0], ))
        self.event_shape = torch.Size((3, ))
        self.device = coeff.device

    @property
    def coefficients(self):
        return [self._coeff[:, l * l:(l + 1)**2].transpose(1, 2).contiguous() for l in range(5)]

    def _s2(self, value):
        y = sph_harm_qm(value.to(self._coeff.dtype))             # [..., B, 25, 2]
        a = self._coeff.sum(dim=2)                                # [B, 25, 2]
        sr = (a[..., 0] * y[..., 0] - a[..., 1] * y[..., 1]).sum(-1)
        si = (a[..., 1] * y[..., 0] + a[..., 0] * y[..., 1]).sum(-1)
        return sr * sr + si * si                                  # [..., B]

    def log_prob_unnormalized(self, value):
        return -self.beta * self._s2(value)

    def prob(self, value):
        if self.beta is not None:
            return torch.exp(self.log_prob(value))
        p = self._s2(value)
        empty = self.empty.reshape((1, ) * (p.dim() - 1) + tuple(self.batch_shape))
        return torch.where(empty, torch.full_like(p, 1 / (4 * math.pi)), p)

    def log_prob(self, value):
        if self.beta is not None:
            return self.log_prob_unnormalized(value) - self.log_z
        return torch.log(self.prob(value).clamp(min=1e-10))

    def _uniform(self, shape):
        # spherical_dists.py:49-61: drawn on the CPU generator like the reference, then moved
        theta = torch.acos(1 - 2 * torch.rand(shape)).to(self.device)
        phi = (2 * math.pi * torch.rand(shape)).to(self.device)
        return torch.stack([torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi), torch.cos(theta)], dim=-1)

    def sample(self, sample_shape=torch.Size()):
        """Rejection sampling against the uniform proposal (spherical_dists.py:116-150, 227-262)."""
        B = self.batch_shape[0]
        num_samples = int(np.prod(sample_shape)) if len(sample_shape) else 1
        log_unif = -math.log(4 * math.pi)
        if self.beta is not None:
            grid = torch.tensor(fibonacci_grid(4096), dtype=self._coeff.dtype, device=self.device).unsqueeze(1)
            log_m = self.log_prob(grid).max(dim=0)[0] - log_unif
            m_value = torch.exp(log_m.clamp(-8, 8))
        else:
            grid = torch.tensor(fibonacci_grid(1024), dtype=self._coeff.dtype, device=self.device).unsqueeze(1)
            m_value = self.prob(grid).max(dim=0)[0] * (4 * math.pi)
            log_m = torch.log(m_value)
        count = min(max(1, int(2 * torch.max(m_value).item())), 1024)
        accepted_t = torch.empty((0, B), dtype=torch.bool, device=self.device)
        candidates_t = torch.empty((0, B, 3), dtype=self._coeff.dtype, device=self.device)
        while bool(torch.any(accepted_t.sum(dim=0) < num_samples)):
            cand = self._uniform((count, B))
            log_thr = self.log_prob(cand) - log_m - log_unif
            u = torch.rand((count, )).unsqueeze(1).to(self.device)
            accepted_t = torch.cat([accepted_t, u < torch.exp(log_thr)], dim=0)
            candidates_t = torch.cat([candidates_t, cand], dim=0)
        samples = torch.stack([candidates_t[:, i][accepted_t[:, i]][:num_samples] for i in range(B)], dim=0)
        return samples.transpose(0, 1).reshape(tuple(sample_shape) + (B, 3)).contiguous()

    def argmax(self, count=None):
        count = count or (128 if self.beta is not None else 256)
        samples = self.sample(torch.Size((count, )))
        score = self.log_prob_unnormalized(samples) if self.beta is not None else self.prob(samples)
        idx = torch.argmax(score, dim=0)
        return torch.gather(samples, 0, idx.view(1, -1, 1).expand(1, -1, 3)).squeeze(0)


class GaussianMixtureModel(D.MixtureSameFamily):
    """gmm.py:8-27."""

    def __init__(self, log_probs, means, stds):
        super().__init__(D.Categorical(logits=log_probs), D.Normal(loc=means, scale=stds))

    def argmax(self, count=128):
        samples = self.sample(torch.Size((count, )))
        idx = torch.argmax(self.log_prob(samples), dim=0).unsqueeze(0)
        return torch.gather(samples, 0, idx).squeeze(0)


class LazyDists:
    """`response['dists']` = [focus, element, distance, so3] distribution objects, built on first access."""

    def __init__(self, agent, fprobs, eprobs, gmm, coeff, log_z, charges):
        self._src = (agent.beta, fprobs, eprobs, gmm, coeff, log_z, charges)
        self._built = None

    def _build(self):
        if self._built is None:
            beta, fprobs, eprobs, gmm, coeff, log_z, charges = self._src
            empty = ~(charges > 0).any(dim=1)
            self._built = [D.Categorical(probs=fprobs.detach()), D.Categorical(probs=eprobs.detach()),
                           GaussianMixtureModel(gmm[:, 0].detach(), gmm[:, 1].detach(), gmm[:, 2].detach()),
                           SO3DistributionView(coeff.detach(), beta, log_z.detach(), empty)]
        return self._built

    def __getitem__(self, i):
        return self._build()[i]

    def __len__(self):
        return 4

    def __iter__(self):
        return iter(self._build())


@torch.no_grad()
def rollout(agent, pos, charges, bags, training: bool):
    """agent.py:229-292 with actions=None: one C-ABI call — the Cormorant body, then ONE kernel that draws focus, element,
    distance and orientation on the device (k_policy_sample: Philox, Categorical inversion, mixture samples, rejection sampling
    on the sphere) and evaluates the chosen action.  The seed comes from torch's global generator (tools/util.py:90-92 seeds it),
    so rollouts are reproducible under set_seeds."""
    seed = int(torch.randint(0, 2**62, (1, ), dtype=torch.int64).item())
    return agent._rollout_raw(pos, charges, bags, mode=1 if training else 2, seed=seed)
