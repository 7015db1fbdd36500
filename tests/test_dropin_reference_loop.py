"""The drop-in contract against the reference's OWN training loop (build container only: needs /root/reference).

The unmodified `molgym.ppo.train` (ppo.py:99-160), `PPOBufferContainer.merge` (buffer_container.py:67-75) and
`DynamicPPOBuffer.get_data` (buffer.py:97-116) drive (a) the reference's own CovariantAC running verbatim on the restated
third-party stand-ins and (b) molgym_b200's CovariantAC — the product's Python class on top of the kernel emulator build
(tests/cusim/emu_agent.py) — from the same parameters, the same synthetic trajectories and the same seeds.  Compared: the
loss info `train` returns, the number of optimizer steps, and the trained policies in function space."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import refrun

needs_reference = pytest.mark.skipif(not refrun.reference_available(), reason='/root/reference not present')


def _fill_container(container, cfg, agent, num_envs, steps, rng):
    """Synthetic trajectories shaped like batch_rollout's (ppo.py:186-209): per step one observation per environment, the
    agent's own value / log-probability of the stored action, random rewards, episodes of three steps."""
    from molgym_b200 import synth
    for t in range(steps):
        obs, n = synth.make_observations(cfg, batch=num_envs, seed=100 + t, start_index=t)
        act = synth.make_actions(cfg, obs, n, seed=200 + t)
        with torch.no_grad():
            pred = agent.step(obs, act)
        terminals = np.full(num_envs, (t % 3) == 2)
        container.store(observations=obs, actions=act, rewards=rng.normal(size=num_envs), next_observations=obs,
                        terminals=terminals, values=pred['v'].detach().cpu().numpy(), logps=pred['logp'].detach().cpu().numpy())
    container.finish_paths(np.zeros(num_envs))


@needs_reference
@pytest.mark.parametrize('fused', [False, True])
def test_reference_train_loop_runs_unchanged_on_the_new_agent(fused):
    refrun.enable(require_reference=True)
    from molgym import ppo as ref_ppo
    from molgym.agents.covariant.agent import CovariantAC as RefCovariantAC
    from molgym.buffer_container import PPOBufferContainer
    from molgym.spaces import ActionSpace, ObservationSpace
    from molgym.tools import util

    from molgym_b200 import ppo as new_ppo
    from molgym_b200 import synth
    from tests.cusim.emu_agent import EmuCovariantAC

    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=32, canvas_size=5, bag={16: 1, 9: 4})
    util.set_seeds(0)
    osp, asp = ObservationSpace(canvas_size=cfg.canvas_size, zs=cfg.zs), ActionSpace(zs=cfg.zs)
    ref = RefCovariantAC(observation_space=osp, action_space=asp, device=torch.device('cpu'), **cfg.agent_kwargs())
    new = EmuCovariantAC(osp, asp, **cfg.agent_kwargs())          # the reference's own space objects, duck-typed
    missing = new.load_state_dict(ref.state_dict())
    assert not missing.missing_keys and not missing.unexpected_keys
    initial = {k: v.clone() for k, v in new.state_dict().items()}

    container = PPOBufferContainer(size=4, gamma=0.99, lam=0.95)
    _fill_container(container, cfg, ref, num_envs=4, steps=6, rng=np.random.default_rng(0))
    data = container.merge().get_data()                             # 24 transitions: minibatches of 10, 10 and a remainder of 4
    assert data['adv'].dtype == np.float64 and data['logp'].dtype == np.float32

    kw = dict(mini_batch_size=10, clip_ratio=0.2, target_kl=0.5, vf_coef=0.5, entropy_coef=0.01, gradient_clip=0.5, max_num_steps=2,
              device=torch.device('cpu'))
    out = {}
    for name, agent in (('ref', ref), ('new', new)):
        optimizer = util.get_optimizer('adam', learning_rate=3e-4, parameters=agent.parameters())
        np.random.seed(7)                                           # get_batch_generator draws its permutation from numpy
        if name == 'new' and fused:
            # the one-line binding of INTEGRATION.md: route train()'s compute_loss to the fused CUDA-graph step
            original, ref_ppo.compute_loss = ref_ppo.compute_loss, new_ppo.compute_loss
            try:
                out[name] = ref_ppo.train(agent, optimizer, data, **kw)
            finally:
                ref_ppo.compute_loss = original
        else:
            new.fused_ppo = False
            out[name] = ref_ppo.train(agent, optimizer, data, **kw)
    assert out['ref']['num_opt_steps'] == out['new']['num_opt_steps'] == 2
    for key in ('policy_loss', 'entropy_loss', 'vf_loss', 'total_loss', 'approx_kl', 'clip_fraction', 'grad_norm'):
        a, b = float(out['ref'][key]), float(out['new'][key])
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (key, a, b)
    with torch.no_grad():
        r, n_ = ref.step(data['obs'], data['act']), new.step(data['obs'], data['act'])
    for key in ('logp', 'ent', 'v'):
        np.testing.assert_allclose(n_[key].numpy(), r[key].numpy(), rtol=2e-4, atol=2e-4, err_msg=key)
    moved = max(float((new.state_dict()[k] - v).abs().max()) for k, v in initial.items())
    assert moved > 1e-4                                             # and the parameters really moved


@needs_reference
def test_reference_rollout_and_checkpoint_helpers_accept_the_new_agent(tmp_path):
    """ppo.py:352-361 toggles `ac.training` as a bare attribute and reads `actions / a / v / logp` from rollout-mode step();
    tools/model_util.py:82-117 pickles the whole module."""
    refrun.enable(require_reference=True)
    from molgym.spaces import ActionSpace, ObservationSpace
    from molgym.tools.util import count_vars, to_numpy

    from molgym_b200 import synth
    from tests.cusim.emu_agent import EmuCovariantAC

    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=32, canvas_size=5, bag={16: 1, 9: 4})
    torch.manual_seed(1)
    osp, asp = ObservationSpace(canvas_size=cfg.canvas_size, zs=cfg.zs), ActionSpace(zs=cfg.zs)
    new = EmuCovariantAC(osp, asp, **cfg.agent_kwargs())
    assert count_vars(new) == sum(int(np.prod(p.shape)) for p in new.parameters())
    obs, n = synth.make_observations(cfg, batch=4)
    for training in (True, False):
        new.training = training
        with torch.no_grad():
            pred = new.step(obs)
        assert len(pred['actions']) == 4 and to_numpy(pred['a']).shape == (4, 6)
        for (element, position), (canvas, bag) in zip(pred['actions'], obs):
            assert bag[element] > 0 and len(position) == 3
    path = tmp_path / 'agent.model'
    torch.save(obj=new, f=str(path))
    loaded = torch.load(f=str(path), weights_only=False)
    act = synth.make_actions(cfg, obs, n)
    with torch.no_grad():
        assert torch.equal(loaded.step(obs, act)['logp'], new.step(obs, act)['logp'])
