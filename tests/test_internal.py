"""Internal-coordinate (SchNet) agent: oracle vs the golden vectors the reference's own code produced, kernel emulator vs
golden / oracle (CPU), and the product path on the GPU."""
import dataclasses
import pickle

import numpy as np
import pytest
import torch

from molgym_b200 import synth
from molgym_b200.agents.internal import zmat
from oracle.molgym_oracle import SchNetOracle, ppo_loss
from oracle.molgym_oracle import internal as oracle_internal
from tests.util_golden import assert_grads_close, assert_outputs_close, golden_grads, golden_observations, golden_state_dict, load_golden

OUT_REL = 1e-5


def _golden():
    g = load_golden('internal_sf6_trained')
    cfg = g['config']
    kw = dict(min_max_distance=tuple(cfg['min_max_distance']), network_width=cfg['network_width'])
    return g, cfg, kw


def test_oracle_reproduces_reference_golden():
    g, cfg, kw = _golden()
    oracle = SchNetOracle(cfg['zs'], cfg['canvas_size'], **kw)
    res = oracle.load_state_dict(golden_state_dict(g), strict=False)
    # golden files hold named_parameters(): buffers (Gaussian offsets, cutoff) and the aliased cfconv.filter_network names are absent
    assert not res.unexpected_keys
    assert all(('.cfconv.filter_network.' in k) or ('distance_expansion' in k) or ('cutoff' in k) for k in res.missing_keys)
    out = oracle.step(golden_observations(g), g['actions'])
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(out[key].detach().numpy(), g[key], rel=2e-6, what=key)
    loss, _ = ppo_loss(out['logp'], out['ent'], out['v'], g['old_logp'], g['adv'], g['ret'], 0.2, 0.5, 0.01)
    assert abs(loss.item() - float(g['loss'])) <= 1e-6
    loss.backward()
    got = {n: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for n, p in oracle.named_parameters()}
    assert_grads_close(got, golden_grads(g))


def test_zmat_matches_oracle_restatement_of_reference():
    """zmat.py:99-133 — 0, 1, 2 and >= 3 existing atoms, both dihedral signs."""
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 3, 6):
        pts = [rng.normal(size=3) * 1.5 for _ in range(n)]
        for focus in range(max(n, 1)):
            for dih in (0.7, -0.7):
                a = zmat.position_atom_helper(pts, focus, 1.3, 1.9, dih)
                b = oracle_internal.position_atom_helper(pts, focus, 1.3, 1.9, dih)
                np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-12)
    with pytest.raises(RuntimeError):
        zmat.position_atom_helper([np.zeros(3)], 2, 1.0, 1.0, 1.0)


def test_batched_molecule_builder_matches_the_per_canvas_statement():
    """zmat.build_molecules (the whole minibatch at once) against build_molecules_loop (one canvas at a time, the reference's
    placement): canvases with 0 .. N atoms incl. full ones, null slots between atoms, both dihedral signs; same errors."""
    rng = np.random.default_rng(3)
    zs, N, B = [0, 1, 6, 8], 5, 40
    obs, act = [], np.zeros((B, 7), dtype=np.float32)
    for b in range(B):
        n = b % (N + 1)
        slots = sorted(rng.choice(N, size=n, replace=False).tolist()) if b % 3 == 0 else list(range(n))   # some canvases with gaps
        canvas = [(0, (0.0, 0.0, 0.0))] * N
        for s_ in slots:
            canvas[s_] = (int(rng.integers(1, len(zs))), tuple((rng.normal(size=3) * 1.5).tolist()))
        obs.append((tuple(canvas), tuple(int(x) for x in rng.integers(0, 3, size=len(zs)))))
        act[b] = [0, rng.integers(0, max(n, 1)), rng.integers(1, len(zs)), rng.uniform(0.9, 2.0), rng.uniform(0.3, 2.8),
                  rng.uniform(0.1, 3.0), rng.integers(0, 2)]
    ref = zmat.build_molecules_loop(obs, act, zs, N)
    got = zmat.build_molecules(obs, act, zs, N)
    assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[2], got[2])
    np.testing.assert_allclose(got[1], ref[1], rtol=0, atol=1e-6)
    bad = act.copy()
    bad[1, 1] = 3            # canvas 1 holds one atom: focus 3 is past it
    for fn in (zmat.build_molecules_loop, zmat.build_molecules):
        with pytest.raises((RuntimeError, IndexError)):
            fn(obs, bad, zs, N)
    worse = [((( 9, (0.0, 0.0, 0.0)), ) * N, (0, 0, 0, 0))]
    for fn in (zmat.build_molecules_loop, zmat.build_molecules):
        with pytest.raises(RuntimeError):
            fn(worse, act[:1], zs, N)


def test_emulator_forward_backward_against_golden():
    from tests.cusim import runner
    g, cfg, kw = _golden()
    sim = runner.CusimInt(cfg['zs'], cfg['canvas_size'], **kw)
    state = golden_state_dict(g)
    flat = sim.flatten(state)
    obs = golden_observations(g)
    numbers, positions, bags = zmat.build_molecules(obs, g['actions'], cfg['zs'], cfg['canvas_size'])
    out = sim.forward(numbers, positions, bags, g['actions'], flat)
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(out[key], g[key], rel=OUT_REL, what=key)
    info, (gl, ge, gv) = runner.ppo_loss(out['logp'], out['ent'], out['v'], g['old_logp'], g['adv'], g['ret'], 0.2, 0.5, 0.01)
    assert abs(info[0] - float(g['loss'])) <= 1e-5
    got = sim.unflatten(sim.backward(gl, ge, gv), {k: tuple(v.shape) for k, v in state.items()})
    assert_grads_close(got, golden_grads(g))


def test_emulator_edge_cases_against_oracle():
    """Empty canvas, one atom, two atoms (auxiliary z-matrix axes), full canvas."""
    from tests.cusim import runner
    cfg = dataclasses.replace(synth.CONFIGS['C1'], canvas_size=4, network_width=32)
    torch.manual_seed(2)
    oracle = SchNetOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    with torch.no_grad():
        for p in oracle.parameters():
            p.add_(0.05 * torch.randn_like(p))
    pad = (cfg.zs.index(0), (0.0, 0.0, 0.0))
    at = lambda k, x, y, z: (k, (x, y, z))
    obs = [((pad, ) * 4, (0, 3, 1)),
           ((at(2, 0.0, 0.0, 0.0), ) + (pad, ) * 3, (0, 3, 0)),
           ((at(2, 0.0, 0.0, 0.0), at(1, 1.4, 0.2, 0.0)) + (pad, ) * 2, (0, 2, 0)),
           ((at(2, 0.0, 0.0, 0.0), at(1, 1.4, 0.2, 0.0), at(1, -0.3, 1.5, 0.4), at(1, 0.1, -0.9, 1.2)), (0, 1, 0))]
    act = np.array([[0, 0, 2, 1.5, 1.2, 0.8, 0], [0, 0, 1, 1.3, 2.0, 0.5, 1], [0, 1, 1, 1.7, 1.1, 2.2, 0], [0, 3, 1, 1.2, 0.9, 1.4, 1]],
                   dtype=np.float32)
    ref = oracle.step(obs, act)
    sim = runner.CusimInt(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    numbers, positions, bags = zmat.build_molecules(obs, act, cfg.zs, cfg.canvas_size)
    out = sim.forward(numbers, positions, bags, act, sim.flatten(oracle.state_dict()))
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(out[key], ref[key].detach().numpy(), rel=OUT_REL, what=key)
    (ref['logp'].sum() - 0.7 * ref['ent'].sum() + 2.0 * ref['v'].sum()).backward()
    ones = np.ones(4, np.float32)
    got = sim.unflatten(sim.backward(ones, -0.7 * ones, 2 * ones), {k: tuple(v.shape) for k, v in oracle.state_dict().items()})
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in oracle.named_parameters()}
    assert_grads_close(got, ref_grads)


# ------------------------------------------------------------------------------------------------------------------
def _agent(zs, canvas_size, **kw):
    from molgym_b200.agents.internal.agent import SchNetAC
    from molgym_b200.spaces import ActionSpace, ObservationSpace
    return SchNetAC(ObservationSpace(canvas_size, zs), ActionSpace(zs), device=torch.device('cuda:0'), **kw)


@pytest.mark.gpu
def test_gpu_golden_step_loss_and_gradients():
    from molgym_b200 import ppo
    g, cfg, kw = _golden()
    agent = _agent(cfg['zs'], cfg['canvas_size'], **kw)
    res = agent.load_state_dict(golden_state_dict(g))
    assert not res.missing_keys and not res.unexpected_keys
    obs = golden_observations(g)
    pred = agent.step(obs, g['actions'])
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(pred[key].detach().cpu().numpy(), g[key], rel=OUT_REL, what=key)
    agent.zero_grad()
    loss, info = ppo.compute_loss(agent, dict(obs=obs, act=g['actions'], logp=g['old_logp'], adv=g['adv'], ret=g['ret']), 0.2, 0.5, 0.01)
    loss.backward()
    assert abs(loss.item() - float(g['loss'])) <= 1e-5
    got = {n: p.grad.detach().cpu().numpy() for n, p in agent.named_parameters()}
    assert_grads_close(got, golden_grads(g))


@pytest.mark.gpu
def test_gpu_c1_config_against_oracle_rollout_and_pickle():
    cfg = synth.CONFIGS['C1']   # SF6, canvas 7, mini_batch 28, width 128: BASELINE.json configs[0]
    torch.manual_seed(5)
    agent = _agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle = SchNetOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle.load_state_dict({k: v.detach().cpu() for k, v in agent.state_dict().items()}, strict=False)
    obs, n = synth.make_observations(cfg)
    act = synth.make_actions(cfg, obs, n)
    ref = oracle.step(obs, act)
    pred = agent.step(obs, act)
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(pred[key].detach().cpu().numpy(), ref[key].detach().numpy(), rel=OUT_REL, what=key)
    w = torch.linspace(-1, 1, len(obs))
    (ref['logp'] * w + ref['v'] - 0.2 * ref['ent']).sum().backward()
    (pred['logp'] * w.cuda() + pred['v'] - 0.2 * pred['ent']).sum().backward()
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in oracle.named_parameters()}
    assert_grads_close({n_: p.grad.detach().cpu().numpy() for n_, p in agent.named_parameters()}, ref_grads)
    for training in (True, False):
        agent.training = training
        with torch.no_grad():
            roll = agent.step(obs[:8])
            again = agent.step(obs[:8], roll['a'].cpu().numpy())
        assert torch.allclose(again['logp'], roll['logp'], rtol=1e-5, atol=1e-5)
        assert len(roll['actions']) == 8
    clone = pickle.loads(pickle.dumps(agent))
    with torch.no_grad():
        assert torch.equal(agent.step(obs, act)['logp'], clone.step(obs, act)['logp'])
