"""The fused optimizer tail (molgym_b200.optim.FlatAdam: gradient norm + clipping + Adam / AMSGrad as two kernels on the flat
buffers) against torch.optim.Adam + torch.nn.utils.clip_grad_norm_ — the sequence molgym/ppo.py:135-146 runs —
on the kernel emulator (CPU) and, marked gpu, on the device; state_dict interchange with torch.optim.Adam; the restated
ppo.train with FlatAdam against the same loop with torch's optimizer."""
import copy

import numpy as np
import pytest
import torch


def _case():
    from tests.cusim import emu_agent
    return emu_agent.make_emu_case()


def _set_grads(agent, seed):
    g = torch.Generator().manual_seed(seed)
    agent.zero_grad()
    agent._attach_grads()
    for p in agent._param_list:     # through the views: the alignment gaps of the flat buffer stay zero, as after a real backward
        p.grad.copy_((torch.randn(p.shape, generator=g) * 3.0).to(p.device))


def _check_against_torch(agent, amsgrad, steps=5, max_norm=0.5):
    from molgym_b200.optim import FlatAdam
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in agent._param_list]
    ref_opt = torch.optim.Adam(ref_params, lr=3e-3, amsgrad=amsgrad)
    opt = FlatAdam(agent, lr=3e-3, amsgrad=amsgrad)
    before = agent._param_version()
    for t in range(steps):
        _set_grads(agent, 10 + t)
        for rp, p in zip(ref_params, agent._param_list):
            rp.grad = p.grad.detach().clone()
        want_norm = torch.nn.utils.clip_grad_norm_(ref_params, max_norm=max_norm)
        ref_opt.step()
        got_norm = opt.step(max_grad_norm=max_norm)
        assert abs(float(got_norm) - float(want_norm)) <= 1e-6 * float(want_norm)
        for rp, p in zip(ref_params, agent._param_list):
            np.testing.assert_allclose(p.detach().cpu().numpy(), rp.detach().cpu().numpy(), rtol=1e-6, atol=3e-7)   # updates are O(lr) = 3e-3
    assert agent._param_version() != before
    return opt, ref_opt, ref_params


@pytest.mark.parametrize('amsgrad', [False, True])
def test_flat_adam_matches_torch_adam_and_clip_on_the_emulator(amsgrad):
    cfg, agent, data = _case()
    opt, ref_opt, ref_params = _check_against_torch(agent, amsgrad)
    # state_dict interchange, both directions (tools/util.py:197-205 builds torch.optim.Adam; checkpoints may hold its state)
    sd, ref_sd = opt.state_dict(), ref_opt.state_dict()
    assert sd['param_groups'][0]['lr'] == ref_sd['param_groups'][0]['lr'] and len(sd['state']) == len(ref_sd['state'])
    for k, st in ref_sd['state'].items():
        for key, val in st.items():
            np.testing.assert_allclose(np.asarray(sd['state'][k][key].cpu()), np.asarray(val.cpu()), rtol=1e-5, atol=1e-8, err_msg=key)
    other = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in agent._param_list], lr=1.0, amsgrad=amsgrad)
    other.load_state_dict(copy.deepcopy(sd))                      # FlatAdam -> torch.optim.Adam
    from molgym_b200.optim import FlatAdam
    again = FlatAdam(agent, lr=1.0, amsgrad=amsgrad)
    again.load_state_dict(copy.deepcopy(ref_sd))                  # torch.optim.Adam -> FlatAdam
    assert again._steps == 5 and again.param_groups[0]['lr'] == 3e-3
    _set_grads(agent, 99)
    for rp, p in zip(ref_params, agent._param_list):
        rp.grad = p.grad.detach().clone()
    ref_opt.step()
    again.step()
    for rp, p in zip(ref_params, agent._param_list):
        np.testing.assert_allclose(p.detach().numpy(), rp.detach().numpy(), rtol=2e-6, atol=3e-7)


def test_restated_train_with_flat_adam_matches_torch_optimizer():
    from molgym_b200 import ppo
    from molgym_b200.optim import FlatAdam
    cfg, agent_a, data = _case()
    _, agent_b, _ = _case()
    kw = dict(mini_batch_size=6, clip_ratio=0.2, target_kl=0.5, vf_coef=0.5, entropy_coef=0.01, gradient_clip=0.5, max_num_steps=2)
    np.random.seed(3)
    info_a = ppo.train(agent_a, torch.optim.Adam(agent_a.parameters(), lr=3e-4), data, **kw)
    np.random.seed(3)
    info_b = ppo.train(agent_b, FlatAdam(agent_b, lr=3e-4), data, **kw)
    assert info_a['num_opt_steps'] == info_b['num_opt_steps'] == 2
    for key in ('policy_loss', 'entropy_loss', 'vf_loss', 'total_loss', 'approx_kl', 'clip_fraction', 'grad_norm'):
        assert abs(info_a[key] - info_b[key]) <= 1e-5 * max(1.0, abs(info_a[key])), key
    with torch.no_grad():
        a, b = agent_a.step(data['obs'], data['act']), agent_b.step(data['obs'], data['act'])
    for key in ('logp', 'ent', 'v'):
        np.testing.assert_allclose(b[key].numpy(), a[key].numpy(), rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('amsgrad', [False, True])
def test_flat_adam_matches_torch_adam_and_clip_on_the_gpu(amsgrad):
    import dataclasses
    from molgym_b200 import _lib, synth
    from molgym_b200.agents.covariant.agent import CovariantAC
    from molgym_b200.spaces import ActionSpace, ObservationSpace
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    torch.manual_seed(0)
    agent = CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=torch.device('cuda:0'), **cfg.agent_kwargs())
    lib = _lib.load()
    before = lib.mgb_launch_count()
    _check_against_torch(agent, amsgrad, steps=5)
    assert lib.mgb_launch_count() - before == 2 * 5        # norm + update per optimizer step
