"""Rollout-mode sampling on the device (k_policy_sample through mgb_cov_rollout).  Distributional checks in the
style of the reference's own (tests/agents/covariant/test_spherical_distr.py:74-101,164-191: mean angles of many samples, atol 0.1;
test_gmm.py:21-25): one observation replicated over the batch, the empirical statistics of the device draws against the distributions
the same step reports (`dists`) and against the torch restatement of the reference's rejection sampler.  Small on the kernel
emulator, large on the GPU."""
import dataclasses

import numpy as np
import pytest
import torch

from molgym_b200 import synth
from tests.util_golden import agent_kwargs_from_config, golden_observations, golden_state_dict, load_golden


def _angles(x):
    x = x / np.linalg.norm(x, axis=-1, keepdims=True)
    return np.arccos(np.clip(x[..., 2], -1, 1)), np.arctan2(x[..., 1], x[..., 0])


def _check_rollout_statistics(make_agent, golden, replicas, calls, atol_freq, atol_angle):
    g = load_golden(golden)
    cfg = g['config']
    agent = make_agent(cfg['zs'], cfg['canvas_size'], **agent_kwargs_from_config(cfg))
    agent.load_state_dict(golden_state_dict(g))
    obs_all = golden_observations(g)
    n_atoms = [sum(1 for lab, _ in canvas if cfg['zs'][lab] != 0) for canvas, _ in obs_all]
    obs = [obs_all[int(np.argmax(n_atoms))]] * replicas          # the fullest canvas of the golden set, replicated
    torch.manual_seed(123)
    agent.training = True
    acts, logps = [], []
    for _ in range(calls):
        with torch.no_grad():
            pred = agent.step(obs)
        acts.append(pred['a'].cpu().numpy())
        logps.append(pred['logp'].cpu().numpy())
    a = np.concatenate(acts)
    assert not np.array_equal(acts[0], acts[1])                    # a new seed per call
    # the stored log-probabilities are those of an evaluate-mode step on the same actions (what ppo.compute_loss recomputes)
    with torch.no_grad():
        again = agent.step(obs, acts[-1])
    np.testing.assert_allclose(again['logp'].cpu().numpy(), logps[-1], rtol=1e-5, atol=1e-5)
    dists = again['dists']
    # focus / element frequencies against the categorical probabilities
    fprobs = dists[0].probs[0].cpu().numpy()
    freq = np.bincount(np.rint(a[:, 0]).astype(int), minlength=len(fprobs)) / len(a)
    assert np.abs(freq - fprobs).max() <= atol_freq, (freq, fprobs)
    assert np.all((a[:, 2] >= 0.001) & np.isfinite(a[:, 2]))
    # conditional statistics for the most frequent (focus, element): distance against the mixture, orientation against the torch sampler
    f0 = int(np.argmax(freq))
    sel = np.rint(a[:, 0]) == f0
    e_freq = np.bincount(np.rint(a[sel, 1]).astype(int), minlength=len(cfg['zs']))
    e0 = int(np.argmax(e_freq))
    sel &= np.rint(a[:, 1]) == e0
    assert sel.sum() >= 20
    probe = a[sel][:1].copy()
    with torch.no_grad():
        cond = agent.step(obs[:1], probe)                            # distributions conditioned on (f0, e0, that distance)
    gmm = cond['dists'][2]
    mean = float((gmm.mixture_distribution.probs * gmm.component_distribution.loc).sum())
    std = float(np.sqrt(((gmm.mixture_distribution.probs * (gmm.component_distribution.scale**2 + gmm.component_distribution.loc**2)).sum()
                         - mean**2).item()))
    assert abs(a[sel, 2].mean() - mean) <= 4 * std / np.sqrt(sel.sum()) + 1e-3
    # orientation: the device sampler against the torch restatement of the reference's rejection sampler, same conditioning distance
    near = sel & (np.abs(a[:, 2] - probe[0, 2]) <= 0.05)
    if near.sum() >= 40:
        so3 = cond['dists'][3]
        ref = so3.sample(torch.Size((int(4 * near.sum()), )))[:, 0].cpu().numpy()
        th_d, ph_d = _angles(a[near, 3:6])
        th_r, ph_r = _angles(ref)
        # the reference's tolerance (atol 0.1 on mean angles) is for its sample counts; here the standard error of the two means
        # (the conditioning on the distance thins the device sample) is added: 3 sigma
        tol = atol_angle + 3.0 * (th_d.std() / np.sqrt(len(th_d)) + th_r.std() / np.sqrt(len(th_r)))
        assert abs(th_d.mean() - th_r.mean()) <= tol, (th_d.mean(), th_r.mean(), tol, len(th_d))
        lp_dv = so3.log_prob(torch.as_tensor(a[near, 3:6], device=so3.device).unsqueeze(1)).cpu().numpy().ravel()
        lp_rv = so3.log_prob(torch.as_tensor(ref, device=so3.device).unsqueeze(1)).cpu().numpy().ravel()
        tol = atol_angle + 3.0 * (lp_dv.std() / np.sqrt(len(lp_dv)) + lp_rv.std() / np.sqrt(len(lp_rv)))
        assert abs(lp_dv.mean() - lp_rv.mean()) <= tol, (lp_dv.mean(), lp_rv.mean(), tol)   # same expected log-density under both samplers
    # greedy mode: argmax of the categoricals, a high-density distance and orientation
    agent.training = False
    with torch.no_grad():
        greedy = agent.step(obs[:8])
    ga = greedy['a'].cpu().numpy()
    assert np.all(np.rint(ga[:, 0]) == int(np.argmax(fprobs)))
    with torch.no_grad():
        gd = agent.step(obs[:8], ga)['dists']
    assert np.all(np.rint(ga[:, 1]) == gd[1].probs.argmax(dim=-1).cpu().numpy())
    lp_greedy = gd[2].log_prob(torch.as_tensor(ga[:, 2], device=gd[2].mixture_distribution.probs.device))
    lp_samples = gd[2].log_prob(gd[2].sample(torch.Size((256, ))))
    assert float(lp_greedy.min()) >= float(torch.quantile(lp_samples.flatten(), 0.5))          # best of 128 mixture samples
    so3 = gd[3]
    lp_o = so3.log_prob(torch.as_tensor(ga[:, 3:6], device=so3.device))                          # [8]
    ref = so3.sample(torch.Size((256, )))
    assert float((lp_o >= torch.quantile(so3.log_prob(ref), 0.75, dim=0)).float().mean()) >= 0.75   # best of >= 128 accepted samples


@pytest.mark.parametrize('golden', ['covariant_sf6_beta', 'covariant_hco_nobeta_trained'])
def test_device_rollout_statistics_on_the_emulator(golden):
    from molgym_b200.spaces import ActionSpace, ObservationSpace
    from tests.cusim.emu_agent import EmuCovariantAC

    def make(zs, canvas_size, **kw):
        return EmuCovariantAC(ObservationSpace(canvas_size, zs), ActionSpace(zs), **kw)
    _check_rollout_statistics(make, golden, replicas=40, calls=3, atol_freq=0.2, atol_angle=0.4)


@pytest.mark.gpu
@pytest.mark.parametrize('golden', ['covariant_sf6_beta', 'covariant_hco_nobeta_trained'])
def test_device_rollout_statistics_on_the_gpu(golden):
    from molgym_b200 import _lib
    from molgym_b200.agents.covariant.agent import CovariantAC
    from molgym_b200.spaces import ActionSpace, ObservationSpace

    def make(zs, canvas_size, **kw):
        return CovariantAC(ObservationSpace(canvas_size, zs), ActionSpace(zs), device=torch.device('cuda:0'), **kw)
    _check_rollout_statistics(make, golden, replicas=1024, calls=4, atol_freq=0.04, atol_angle=0.05)
    # one rollout step = one C-ABI call, no host round trip between the sub-actions
    lib = _lib.load()
    g = load_golden(golden)
    cfg = g['config']
    agent = make(cfg['zs'], cfg['canvas_size'], **agent_kwargs_from_config(cfg))
    obs = golden_observations(g)[:10]
    with torch.no_grad():
        agent.step(obs)
        before = lib.mgb_launch_count()
        agent.step(obs)
    assert lib.mgb_launch_count() - before <= 24      # the body (20 launches for three CG levels) + ONE sampling kernel
