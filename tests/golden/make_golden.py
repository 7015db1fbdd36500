"""Generate golden vectors by running the REFERENCE'S OWN code verbatim (PYTHONPATH=/root/reference) on the
restated third-party stand-ins (oracle/thirdparty).  Build-container only; the .npz files it writes are
committed so the GPU box never needs /root/reference.

    python tests/golden/make_golden.py
"""
import dataclasses
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refrun  # noqa: E402

refrun.enable(require_reference=True)

from molgym import ppo  # noqa: E402
from molgym.agents.covariant.agent import CovariantAC  # noqa: E402
from molgym.agents.internal.agent import SchNetAC  # noqa: E402
from molgym.spaces import ActionSpace, ObservationSpace  # noqa: E402
from molgym.tools import util  # noqa: E402

from molgym_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CLIP, VF, ENT = 0.2, 0.5, 0.01  # arg_parser.py:84-86


def obs_to_arrays(observations):
    labels = np.array([[item[0] for item in canvas] for canvas, _ in observations], dtype=np.int32)
    xyz = np.array([[item[1] for item in canvas] for canvas, _ in observations], dtype=np.float64)
    bags = np.array([bag for _, bag in observations], dtype=np.int64)
    return labels, xyz, bags


def covariant_case(name, cfg, batch, pre_steps):
    util.set_seeds(0)
    osp = ObservationSpace(canvas_size=cfg.canvas_size, zs=cfg.zs)
    agent = CovariantAC(observation_space=osp, action_space=ActionSpace(zs=cfg.zs), device=torch.device('cpu'),
                        **cfg.agent_kwargs())
    obs, n_atoms = synth.make_observations(cfg, batch=batch)
    act = synth.make_actions(cfg, obs, n_atoms)
    with torch.no_grad():
        logp0 = agent.step(obs, act)['logp'].numpy()
    old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
    data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)
    # optionally move the parameters away from their initial values with the reference's own update rule
    opt = torch.optim.Adam(agent.parameters(), lr=3e-3)
    for _ in range(pre_steps):
        opt.zero_grad()
        loss, _ = ppo.compute_loss(agent, data, CLIP, VF, ENT)
        loss.backward()
        opt.step()
    agent.zero_grad()
    pred = agent.step(obs, act)
    loss, info = ppo.compute_loss(agent, data, CLIP, VF, ENT)
    loss.backward()
    so3 = pred['dists'][-1]
    labels, xyz, bags = obs_to_arrays(obs)
    out = dict(labels=labels, xyz=xyz, bags=bags, actions=act, old_logp=old_logp, adv=adv, ret=ret,
               logp=pred['logp'].detach().numpy(), ent=pred['ent'].detach().numpy(), v=pred['v'].detach().numpy(),
               focus_probs=pred['dists'][0].probs.detach().numpy(),
               element_probs=pred['dists'][1].probs.detach().numpy(),
               loss=np.array(loss.item()), **{'info_' + k: np.array(v) for k, v in info.items()})
    for ell, part in enumerate(so3.coefficients):
        out[f'coeff_{ell}'] = part.detach().numpy()
    if hasattr(so3, 'log_z'):
        out['log_z'] = so3.log_z.detach().numpy()
    for pname, p in agent.named_parameters():
        out['param/' + pname] = p.detach().numpy()
        out['grad/' + pname] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    out['config_json'] = np.array(repr(dataclasses.asdict(cfg)))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'loss', loss.item(), 'logp[:3]', out['logp'][:3], 'params', sum(p.numel() for p in agent.parameters()))


def internal_case(name, cfg, batch, pre_steps):
    util.set_seeds(0)
    osp = ObservationSpace(canvas_size=cfg.canvas_size, zs=cfg.zs)
    agent = SchNetAC(observation_space=osp, action_space=ActionSpace(zs=cfg.zs), device=torch.device('cpu'), **cfg.agent_kwargs())
    obs, n_atoms = synth.make_observations(cfg, batch=batch)
    act = synth.make_actions(cfg, obs, n_atoms)
    with torch.no_grad():
        logp0 = agent.step(obs, act)['logp'].numpy()
    old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
    data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)
    opt = torch.optim.Adam(agent.parameters(), lr=3e-3)
    for _ in range(pre_steps):   # moves the zero-initialised biases away from zero with the reference's own update rule
        opt.zero_grad()
        loss, _ = ppo.compute_loss(agent, data, CLIP, VF, ENT)
        loss.backward()
        opt.step()
    agent.zero_grad()
    pred = agent.step(obs, act)
    loss, info = ppo.compute_loss(agent, data, CLIP, VF, ENT)
    loss.backward()
    labels, xyz, bags = obs_to_arrays(obs)
    out = dict(labels=labels, xyz=xyz, bags=bags, actions=act, old_logp=old_logp, adv=adv, ret=ret,
               logp=pred['logp'].detach().numpy(), ent=pred['ent'].detach().numpy(), v=pred['v'].detach().numpy(),
               loss=np.array(loss.item()), **{'info_' + k: np.array(v) for k, v in info.items()})
    for pname, p in agent.named_parameters():
        out['param/' + pname] = p.detach().numpy()
        out['grad/' + pname] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    out['config_json'] = np.array(repr(dataclasses.asdict(cfg)))
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'loss', loss.item(), 'logp[:3]', out['logp'][:3], 'params', sum(p.numel() for p in agent.parameters()))


if __name__ == '__main__':
    if 'internal' in sys.argv or len(sys.argv) == 1:
        c1 = dataclasses.replace(synth.CONFIGS['C1'], network_width=64)
        internal_case('internal_sf6_trained', c1, batch=14, pre_steps=2)
        if 'internal' in sys.argv:
            sys.exit(0)
    c2 = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    covariant_case('covariant_sf6_beta', c2, batch=14, pre_steps=0)
    small = dataclasses.replace(synth.CONFIGS['C2'], name='small-HCO', zs=[0, 1, 6, 8], canvas_size=5,
                                bag={6: 1, 1: 3, 8: 1}, min_max_distance=(0.9, 1.8), bag_scale=3, beta=None,
                                network_width=64, seed=77)
    covariant_case('covariant_hco_nobeta_trained', small, batch=10, pre_steps=3)
