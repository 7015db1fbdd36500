"""Kernel-logic parity on the CPU: the product's .cu sources compiled against the kernel emulator (tests/cusim) and
driven through the same C ABI, compared with (a) the golden vectors the reference's own code produced and (b) the
oracle on fresh canvases.  The GPU parity tests proper are in test_gpu_parity.py; these catch indexing / barrier /
reduction / hand-written-backward mistakes before GPU time is spent."""
import dataclasses

import numpy as np
import pytest
import torch

from molgym_b200 import synth
from oracle.molgym_oracle import CovariantOracle, pack_observations, ppo_loss
from tests.cusim import runner
from tests.util_golden import (agent_kwargs_from_config, assert_grads_close, assert_outputs_close, golden_grads,
                               golden_observations, golden_state_dict, load_golden)

OUT_REL = 1e-5   # outputs: abs(x - ref) <= 1e-5 * max(abs(ref), 1e-3)  


@pytest.mark.parametrize('name', ['covariant_sf6_beta', 'covariant_hco_nobeta_trained'])
def test_golden_forward_backward(name):
    g = load_golden(name)
    cfg = g['config']
    sim = runner.CusimCov(cfg['zs'], cfg['canvas_size'], **agent_kwargs_from_config(cfg))
    state = golden_state_dict(g)
    flat = sim.flatten(state)
    pos, charges, bags = pack_observations(golden_observations(g), cfg['zs'], cfg['canvas_size'])
    out = sim.forward(pos, charges, bags, g['actions'], flat)
    for key in ('logp', 'ent', 'v', 'focus_probs', 'element_probs'):
        assert_outputs_close(out[key], g[key], rel=OUT_REL, what=key)
    for ell in range(5):
        got = np.transpose(out['coefficients'][:, ell * ell:(ell + 1)**2], (0, 2, 1, 3))
        assert_outputs_close(got, g[f'coeff_{ell}'], rel=OUT_REL, floor=1.0, what=f'coeff_{ell}')  # unit-normalised vector
    if 'log_z' in g:
        assert_outputs_close(out['log_z'], g['log_z'], rel=OUT_REL, what='log_z')
    info, (g_logp, g_ent, g_v) = runner.ppo_loss(out['logp'], out['ent'], out['v'], g['old_logp'], g['adv'], g['ret'], 0.2, 0.5, 0.01)
    assert abs(info[0] - float(g['loss'])) <= 1e-5 * max(1.0, abs(float(g['loss'])))   # terms of O(|adv|) = O(1) cancel
    for idx, key in ((1, 'policy_loss'), (2, 'entropy_loss'), (3, 'vf_loss'), (4, 'approx_kl'), (5, 'clip_fraction')):
        assert abs(info[idx] - float(g['info_' + key])) <= 1e-5, key
    grad = sim.backward(g_logp, g_ent, g_v)
    got = sim.unflatten(grad, {k: tuple(v.shape) for k, v in state.items()})
    assert_grads_close(got, golden_grads(g))


@pytest.mark.parametrize('levels,beta,zs,canvas,edge_mode', [(1, -10.0, [0, 9, 16], 7, None), (2, None, [0, 1, 6, 7, 8], 6, '0'),
                                                             (2, None, [0, 1, 6, 7, 8], 6, '1'), (2, None, [0, 1, 6, 7, 8], 6, '2'),
                                                             (2, -5.0, [0, 1, 6, 7, 8, 9], 5, None)])   # 24 output channels: two passes of the weight-gradient kernel, CO = 32 row mix
def test_fresh_canvases_against_oracle(levels, beta, zs, canvas, edge_mode, monkeypatch):
    if edge_mode is not None:   # the three decompositions of the per-pair edge kernels (plan.cuh: edge_mode_override)
        monkeypatch.setenv('MGB_EDGE_MODE', edge_mode)
        if edge_mode == '0':    # ... and the large-minibatch atom path (combined atom kernels, tiled InputLinear gradient)
            monkeypatch.setenv('MGB_SMALL_ATOMS', '0')
            monkeypatch.setenv('MGB_MIX_TC', '1')   # ... and the tensor-core channel mix (both directions)
            monkeypatch.setenv('MGB_LARGE_ATOMS', '1')   # ... and the rest of the large-minibatch choices (512-thread dcat mix, two-pass mix weight gradient)
    cfg = dataclasses.replace(synth.CONFIGS['C3'], zs=zs, canvas_size=canvas, network_width=32, num_cg_levels=levels, beta=beta,
                              bag={z: 2 for z in zs if z}, bag_scale=4, seed=levels)
    torch.manual_seed(levels)
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    B = 7
    obs, n = synth.make_observations(cfg, batch=B)
    act = synth.make_actions(cfg, obs, n)
    pos, charges, bags = pack_observations(obs, cfg.zs, cfg.canvas_size)
    ref = oracle.evaluate(pos, charges, bags, act)
    sim = runner.CusimCov(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    out = sim.forward(pos, charges, bags, act, sim.flatten(oracle.state_dict()))
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(out[key], ref[key].detach().numpy(), rel=OUT_REL, what=key)
    for ell in range(5):
        r = ref['covariats'][ell].detach().numpy()
        got = np.transpose(out['covariats'][:, :, ell * ell:(ell + 1)**2], (0, 1, 3, 2, 4))
        assert_outputs_close(got, r, rel=OUT_REL, floor=max(np.abs(r).max(), 1e-6), what=f'covariats_{ell}')  # norm-wise: components mix under rotation
    rng = np.random.default_rng(1)
    gl, ge, gv = (rng.normal(size=B).astype(np.float32) for _ in range(3))
    (ref['logp'] * torch.tensor(gl) + ref['ent'] * torch.tensor(ge) + ref['v'] * torch.tensor(gv)).sum().backward()
    grad = sim.unflatten(sim.backward(gl, ge, gv), {k: tuple(v.shape) for k, v in oracle.state_dict().items()})
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(p.shape, np.float32)) for k, p in oracle.named_parameters()}
    assert_grads_close(grad, ref_grads)


@pytest.mark.parametrize('which,batch,start,large', [('C5', 3, 37, True), ('C5', 3, 24, False), ('C4', 4, 18, True)])
def test_high_occupancy_canvases_against_oracle(which, batch, start, large, monkeypatch):
    """Canvases with 18-21 (N = 22) and 24-39 (N = 40) atoms: three to five neighbour chunks per atom in the atom kernels, the
    widest shared-memory carve; `large` forces the decomposition the full-size minibatches of C3-C5 run on (thread per pair
    edge kernels, combined atom kernels, single-kernel policy backward, tiled InputLinear gradient)."""
    if large:
        monkeypatch.setenv('MGB_EDGE_MODE', '0')
        monkeypatch.setenv('MGB_SMALL_ATOMS', '0')
        if which == 'C4':   # the tensor-core channel mix in both directions; the other cases take the default large path (tensor-core forward, 512-thread FFMA backward)
            monkeypatch.setenv('MGB_MIX_TC', '1')
        monkeypatch.setenv('MGB_LARGE_ATOMS', '1')   # ... and the rest of the large-minibatch choices (512-thread dcat mix, two-pass mix weight gradient)
    cfg = dataclasses.replace(synth.CONFIGS[which], network_width=32)
    torch.manual_seed(8)
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    # the float64 oracle is the arbiter here: with ~40 atoms per canvas the float32 torch formulation itself is
    # 5e-4 off in some gradients (phi_focus.layers.0.weight), the kernels are not
    arbiter = CovariantOracle(cfg.zs, cfg.canvas_size, dtype=torch.float64, **cfg.agent_kwargs())
    arbiter.load_state_dict({k: v.double() for k, v in oracle.state_dict().items()})
    obs, n = synth.make_observations(cfg, batch=batch, start_index=start)
    assert n.min() >= start and n.max() == min(start + batch - 1, cfg.canvas_size - 1)
    act = synth.make_actions(cfg, obs, n)
    pos, charges, bags = pack_observations(obs, cfg.zs, cfg.canvas_size)
    ref = arbiter.evaluate(pos, charges, bags, act)
    sim = runner.CusimCov(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    out = sim.forward(pos, charges, bags, act, sim.flatten(oracle.state_dict()))
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(out[key], ref[key].detach().numpy(), rel=OUT_REL, what=key)
    rng = np.random.default_rng(4)
    gl, ge, gv = (rng.normal(size=batch).astype(np.float32) for _ in range(3))
    (ref['logp'] * torch.tensor(gl, dtype=torch.float64) + ref['ent'] * torch.tensor(ge, dtype=torch.float64)
     + ref['v'] * torch.tensor(gv, dtype=torch.float64)).sum().backward()
    grad = sim.unflatten(sim.backward(gl, ge, gv), {k: tuple(v.shape) for k, v in oracle.state_dict().items()})
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(p.shape, np.float64)) for k, p in arbiter.named_parameters()}
    assert_grads_close(grad, ref_grads)


@pytest.mark.parametrize('hidden,edge_mode', [(6, None), (5, '0')])
def test_odd_hyperparameters_take_the_generic_paths(hidden, edge_mode, monkeypatch):
    """6 hidden channels (run-time channel stride instead of the compile-time 10, zero-padded edge register tiles), 3 channels
    per element, network width 30 (not a multiple of 4: the row MLPs fall back from the bulk-copy kernels), 2 Gaussians.
    5 hidden channels with the thread-per-pair edge kernels: odd channel counts take 8-byte instead of 16-byte row accesses."""
    if edge_mode is not None:
        monkeypatch.setenv('MGB_EDGE_MODE', edge_mode)
        monkeypatch.setenv('MGB_SMALL_ATOMS', '0')
        monkeypatch.setenv('MGB_LARGE_ATOMS', '1')
    zs = [0, 1, 8]
    cfg = dataclasses.replace(synth.CONFIGS['C3'], zs=zs, canvas_size=5, network_width=30, num_cg_levels=2, beta=-3.0,
                              num_channels_hidden=hidden, num_channels_per_element=3, num_gaussians=2, bag={1: 3, 8: 2}, bag_scale=4, seed=5)
    torch.manual_seed(4)
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    B = 6
    obs, n = synth.make_observations(cfg, batch=B)
    act = synth.make_actions(cfg, obs, n)
    pos, charges, bags = pack_observations(obs, cfg.zs, cfg.canvas_size)
    ref = oracle.evaluate(pos, charges, bags, act)
    sim = runner.CusimCov(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    out = sim.forward(pos, charges, bags, act, sim.flatten(oracle.state_dict()))
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(out[key], ref[key].detach().numpy(), rel=OUT_REL, what=key)
    rng = np.random.default_rng(2)
    gl, ge, gv = (rng.normal(size=B).astype(np.float32) for _ in range(3))
    (ref['logp'] * torch.tensor(gl) + ref['ent'] * torch.tensor(ge) + ref['v'] * torch.tensor(gv)).sum().backward()
    grad = sim.unflatten(sim.backward(gl, ge, gv), {k: tuple(v.shape) for k, v in oracle.state_dict().items()})
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(p.shape, np.float32)) for k, p in oracle.named_parameters()}
    assert_grads_close(grad, ref_grads)


def test_edge_canvases_empty_single_full():
    """Empty canvas (bias-only logits, uniform orientation), one atom, completely filled canvas."""
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=32, canvas_size=4, bag={16: 1, 9: 4})
    torch.manual_seed(5)
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    null = cfg.zs.index(0)
    pad = (null, (0.0, 0.0, 0.0))
    full = tuple((1 + (i % 2), (0.9 * i, 0.3 * (i % 2), -0.2 * i)) for i in range(4))
    obs = [((pad, ) * 4, (0, 4, 1)), (((2, (0.0, 0.0, 0.0)), ) + (pad, ) * 3, (0, 4, 0)), (full, (0, 1, 0))]
    act = np.array([[0, 2, 1.5, 0, 0, 1], [0, 1, 1.3, 0.6, 0.0, 0.8], [3, 1, 1.9, -1, 0, 0]], dtype=np.float32)
    pos, charges, bags = pack_observations(obs, cfg.zs, cfg.canvas_size)
    ref = oracle.evaluate(pos, charges, bags, act)
    sim = runner.CusimCov(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    out = sim.forward(pos, charges, bags, act, sim.flatten(oracle.state_dict()))
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(out[key], ref[key].detach().numpy(), rel=OUT_REL, what=key)
    (ref['logp'].sum() + 0.5 * ref['ent'].sum() - ref['v'].sum()).backward()
    ones = np.ones(3, np.float32)
    grad = sim.unflatten(sim.backward(ones, 0.5 * ones, -ones), {k: tuple(v.shape) for k, v in oracle.state_dict().items()})
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(p.shape, np.float32)) for k, p in oracle.named_parameters()}
    assert_grads_close(grad, ref_grads)


def test_ppo_loss_kernel_matches_oracle_including_clipped_and_tied_branches():
    rng = np.random.default_rng(3)
    B = 257
    logp = rng.normal(size=B).astype(np.float32)
    old = (logp + rng.normal(scale=0.25, size=B)).astype(np.float32)
    old[:5] = logp[:5]   # ratio == 1 exactly: torch.min tie
    ent, v = rng.normal(size=B).astype(np.float32), rng.normal(size=B).astype(np.float32)
    adv, ret = rng.normal(size=B), rng.normal(size=B)
    tl, te, tv = (torch.tensor(x, requires_grad=True) for x in (logp, ent, v))
    loss, info = ppo_loss(tl, te, tv, old, adv, ret, 0.2, 0.5, 0.01)
    loss.backward()
    kinfo, (gl, ge, gv) = runner.ppo_loss(logp, ent, v, old, adv, ret, 0.2, 0.5, 0.01)
    assert abs(kinfo[0] - loss.item()) < 1e-12 + 1e-9 * abs(loss.item())
    assert abs(kinfo[4] - info['approx_kl']) < 1e-6 and abs(kinfo[5] - info['clip_fraction']) < 1e-7
    np.testing.assert_allclose(gl, tl.grad.numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(ge, te.grad.numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(gv, tv.grad.numpy(), rtol=1e-6, atol=1e-9)


def test_scale_accumulate_kernel():
    """.grad (+)= cotangent * scratch gradient of the fused PPO step (mgb_scale_accumulate): float64 / float32 device scalar,
    overwrite / accumulate, lengths that are not a multiple of the vector width."""
    rng = np.random.default_rng(3)
    for n in (1, 7, 64, 1003):
        dst, src = rng.normal(size=n).astype(np.float32), rng.normal(size=n).astype(np.float32)
        for scale in (np.float64(0.25), np.float32(-1.5)):
            got = runner.scale_accumulate(dst, src, scale, accumulate=False)
            np.testing.assert_array_equal(got, (np.float32(scale) * src).astype(np.float32))
            got = runner.scale_accumulate(dst, src, scale, accumulate=True)
            ref = np.float32(scale) * src.astype(np.float64) + dst   # one fused multiply-add per element
            np.testing.assert_allclose(got, ref.astype(np.float32), rtol=1e-6, atol=1e-7)


def test_packer_matches_oracle_and_rejects_bad_labels():
    cfg = synth.CONFIGS['C3']
    obs, _ = synth.make_observations(cfg, batch=9)
    # a null item in the middle of the canvas must be dropped and the rest compacted (spaces.py:55-61)
    canvas, bag = obs[5]
    canvas = (canvas[0], (cfg.zs.index(0), (9.0, 9.0, 9.0))) + canvas[1:-1]
    obs[5] = (canvas, bag)
    labels = np.array([[it[0] for it in c] for c, _ in obs], np.int32)
    xyz = np.array([[it[1] for it in c] for c, _ in obs], np.float64)
    pos, charges = runner.pack(cfg.zs, cfg.canvas_size, labels, xyz)
    rpos, rcharges, _ = pack_observations(obs, cfg.zs, cfg.canvas_size)
    assert np.array_equal(pos, rpos) and np.array_equal(charges, rcharges)
    labels[0, 0] = -1
    with pytest.raises(RuntimeError):
        runner.pack(cfg.zs, cfg.canvas_size, labels, xyz)
