"""GPU parity tests (run on the B200 box with -m gpu): the product path — CovariantAC.step through the C ABI and the
sm_100a kernels — against the committed golden vectors (reference's own code), the oracle on fresh canvases, and
size-independent properties at BASELINE.json's full minibatch sizes."""
import dataclasses
import pickle

import numpy as np
import pytest
import torch

from molgym_b200 import synth
from tests.util_golden import (agent_kwargs_from_config, assert_grads_close, assert_outputs_close, golden_grads,
                               golden_observations, golden_state_dict, load_golden)

pytestmark = pytest.mark.gpu
OUT_REL = 1e-5   # abs(x - ref) <= 1e-5 * max(abs(ref), 1e-3)  (north_star: logits/value within 1e-5 relative)


def make_agent(cfg_zs, canvas_size, **kw):
    from molgym_b200.agents.covariant.agent import CovariantAC
    from molgym_b200.spaces import ActionSpace, ObservationSpace
    return CovariantAC(ObservationSpace(canvas_size, cfg_zs), ActionSpace(cfg_zs), device=torch.device('cuda:0'), **kw)


def grads_of(agent):
    return {n: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
            for n, p in agent.named_parameters()}


def test_native_library_is_loaded_and_is_the_cuda_build():
    from molgym_b200 import _lib
    lib = _lib.load()
    assert lib.mgb_is_cuda_build() == 1
    maps = open('/proc/self/maps').read()
    assert 'libmolgym_b200.so' in maps


@pytest.mark.parametrize('name', ['covariant_sf6_beta', 'covariant_hco_nobeta_trained'])
def test_golden_step_loss_and_gradients(name):
    from molgym_b200 import ppo
    g = load_golden(name)
    cfg = g['config']
    agent = make_agent(cfg['zs'], cfg['canvas_size'], **agent_kwargs_from_config(cfg))
    missing = agent.load_state_dict(golden_state_dict(g))
    assert not missing.missing_keys and not missing.unexpected_keys
    obs = golden_observations(g)
    pred = agent.step(obs, g['actions'])
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(pred[key].detach().cpu().numpy(), g[key], rel=OUT_REL, what=key)
    assert_outputs_close(pred['dists'][0].probs.cpu().numpy(), g['focus_probs'], rel=OUT_REL, what='focus_probs')
    assert_outputs_close(pred['dists'][1].probs.cpu().numpy(), g['element_probs'], rel=OUT_REL, what='element_probs')
    for ell, part in enumerate(pred['dists'][3].coefficients):
        assert_outputs_close(part.cpu().numpy(), g[f'coeff_{ell}'], rel=OUT_REL, floor=1.0, what=f'coeff_{ell}')
    agent.zero_grad()
    loss, info = ppo.compute_loss(agent, dict(obs=obs, act=g['actions'], logp=g['old_logp'], adv=g['adv'], ret=g['ret']), 0.2, 0.5, 0.01)
    loss.backward()
    assert abs(loss.item() - float(g['loss'])) <= 1e-5
    for key in ('policy_loss', 'vf_loss', 'entropy_loss', 'approx_kl', 'clip_fraction'):
        assert abs(info[key] - float(g['info_' + key])) <= 1e-5, key
    assert_grads_close(grads_of(agent), golden_grads(g))


@pytest.mark.parametrize('which,batch', [('C2', 140), ('C3', 96), ('C4', 30), ('C5', 24)])
def test_fresh_canvases_against_oracle(which, batch):
    from oracle.molgym_oracle import CovariantOracle, ppo_loss
    from molgym_b200 import ppo
    cfg = synth.CONFIGS[which]
    torch.manual_seed(11)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle.load_state_dict({k: v.detach().cpu() for k, v in agent.state_dict().items()})
    obs, n = synth.make_observations(cfg, batch=batch)
    act = synth.make_actions(cfg, obs, n)
    ref = oracle.step(obs, act)
    old_logp, adv, ret = synth.make_ppo_targets(cfg, ref['logp'].detach().numpy())
    loss, info = ppo.compute_loss(agent, dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret), 0.2, 0.5, 0.01)
    loss.backward()
    pred = agent.step(obs, act)
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(pred[key].detach().cpu().numpy(), ref[key].detach().numpy(), rel=OUT_REL, what=key)
    ref_loss, _ = ppo_loss(ref['logp'], ref['ent'], ref['v'], old_logp, adv, ret, 0.2, 0.5, 0.01)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in oracle.named_parameters()}
    assert_grads_close(grads_of(agent), ref_grads)


@pytest.mark.parametrize('which,batch,start,large', [('C5', 16, 24, True), ('C5', 16, 24, False), ('C4', 30, 0, True), ('C5', 24, 0, True)])
def test_large_decomposition_and_high_occupancy_against_oracle(which, batch, start, large, monkeypatch):
    """The decomposition the full-size C3-C5 minibatches run on (MGB_EDGE_MODE=0 + MGB_SMALL_ATOMS=0: thread-per-pair edge
    kernels, combined atom kernels, single-kernel policy backward, tiled InputLinear gradient) at N = 22 and N = 40, and the
    C5 occupancies 24..39 (three to five neighbour chunks per atom) that a batch counted from zero never reaches."""
    from oracle.molgym_oracle import CovariantOracle, ppo_loss
    from molgym_b200 import ppo
    if large:
        monkeypatch.setenv('MGB_EDGE_MODE', '0')
        monkeypatch.setenv('MGB_SMALL_ATOMS', '0')
        if which == 'C4':   # the tensor-core channel mix in both directions; the other cases take the default large path (tensor-core forward, 512-thread FFMA backward)
            monkeypatch.setenv('MGB_MIX_TC', '1')
        monkeypatch.setenv('MGB_LARGE_ATOMS', '1')   # ... and the rest of the large-minibatch choices (512-thread dcat mix, two-pass mix weight gradient)
    cfg = synth.CONFIGS[which]
    torch.manual_seed(13)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    # the float64 oracle is the arbiter: at 20-40 atoms per canvas the float32 torch formulation itself is a few
    # 1e-4 off in some gradients, the kernels are not
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, dtype=torch.float64, **cfg.agent_kwargs())
    oracle.load_state_dict({k: v.detach().cpu().double() for k, v in agent.state_dict().items()})
    obs, n = synth.make_observations(cfg, batch=batch, start_index=start)
    if start:
        assert n.min() >= start and n.max() == cfg.canvas_size - 1
    act = synth.make_actions(cfg, obs, n)
    ref = oracle.step(obs, act)
    old_logp, adv, ret = synth.make_ppo_targets(cfg, ref['logp'].detach().numpy())
    ref_loss, _ = ppo_loss(ref['logp'], ref['ent'], ref['v'], old_logp, adv, ret, 0.2, 0.5, 0.01)
    ref_loss.backward()
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape))) for k, p in oracle.named_parameters()}
    for fused in (True, False):
        agent.fused_ppo = fused
        agent.zero_grad()
        loss, _ = ppo.compute_loss(agent, dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret), 0.2, 0.5, 0.01)
        loss.backward()
        assert abs(loss.item() - ref_loss.item()) <= 1e-5 * max(1.0, abs(ref_loss.item()))
        assert_grads_close(grads_of(agent), ref_grads)
    with torch.no_grad():
        pred = agent.step(obs, act)
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(pred[key].cpu().numpy(), ref[key].detach().numpy(), rel=OUT_REL, what=key)


def test_six_species_24_output_channels_against_oracle():
    """Z = 6 species x 4 channels = 24 output channels: the widest register tile of the row mix (CO = 32) and two channel
    passes of the atom-mix weight gradient, paths the benchmark configurations (Z <= 5) never take."""
    from oracle.molgym_oracle import CovariantOracle, ppo_loss
    from molgym_b200 import ppo
    zs = [0, 1, 6, 7, 8, 9]
    cfg = dataclasses.replace(synth.CONFIGS['C3'], zs=zs, canvas_size=6, network_width=64, bag={z: 2 for z in zs if z}, bag_scale=4, seed=3)
    torch.manual_seed(9)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle.load_state_dict({k: v.detach().cpu() for k, v in agent.state_dict().items()})
    obs, n = synth.make_observations(cfg, batch=20)
    act = synth.make_actions(cfg, obs, n)
    ref = oracle.step(obs, act)
    old_logp, adv, ret = synth.make_ppo_targets(cfg, ref['logp'].detach().numpy())
    loss, _ = ppo.compute_loss(agent, dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret), 0.2, 0.5, 0.01)
    loss.backward()
    ref_loss, _ = ppo_loss(ref['logp'], ref['ent'], ref['v'], old_logp, adv, ret, 0.2, 0.5, 0.01)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in oracle.named_parameters()}
    assert_grads_close(grads_of(agent), ref_grads)


@pytest.mark.parametrize('hidden,edge_mode', [(6, None), (5, '0')])
def test_odd_hyperparameters_against_oracle(hidden, edge_mode, monkeypatch):
    """6 hidden channels (run-time channel stride, zero-padded edge tiles), 3 channels per element (9 output channels), width 30
    (the row MLPs fall back from the bulk-copy kernels; unaligned bulk copies fall back to plain loads), 2 Gaussians.
    5 hidden channels with the thread-per-pair edge kernels: odd channel counts take 8-byte instead of 16-byte row accesses."""
    from oracle.molgym_oracle import CovariantOracle, ppo_loss
    from molgym_b200 import ppo
    if edge_mode is not None:
        monkeypatch.setenv('MGB_EDGE_MODE', edge_mode)
        monkeypatch.setenv('MGB_SMALL_ATOMS', '0')
        monkeypatch.setenv('MGB_LARGE_ATOMS', '1')
    zs = [0, 1, 8]
    cfg = dataclasses.replace(synth.CONFIGS['C3'], zs=zs, canvas_size=5, network_width=30, num_cg_levels=2, beta=-3.0,
                              num_channels_hidden=hidden, num_channels_per_element=3, num_gaussians=2, bag={1: 3, 8: 2}, bag_scale=4, seed=5)
    torch.manual_seed(4)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle.load_state_dict({k: v.detach().cpu() for k, v in agent.state_dict().items()})
    obs, n = synth.make_observations(cfg, batch=25)
    act = synth.make_actions(cfg, obs, n)
    ref = oracle.step(obs, act)
    old_logp, adv, ret = synth.make_ppo_targets(cfg, ref['logp'].detach().numpy())
    for fused in (True, False):
        agent.fused_ppo = fused
        agent.zero_grad()
        loss, _ = ppo.compute_loss(agent, dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret), 0.2, 0.5, 0.01)
        loss.backward()
        if fused:
            ref_loss, _ = ppo_loss(ref['logp'], ref['ent'], ref['v'], old_logp, adv, ret, 0.2, 0.5, 0.01)
            ref_loss.backward()
        assert abs(loss.item() - ref_loss.item()) <= 1e-5
        ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in oracle.named_parameters()}
        assert_grads_close(grads_of(agent), ref_grads)


@pytest.mark.parametrize('mode', ['0', '1', '2'])
def test_edge_kernel_decompositions_agree_with_oracle(mode, monkeypatch):
    """The per-pair edge kernels exist in three decompositions picked by minibatch size (thread per pair, per (pair, ell),
    five threads per (pair, ell)); MGB_EDGE_MODE forces each of them on the same canvases.  Mode 0 also forces the
    large-minibatch atom path (MGB_SMALL_ATOMS=0), which the small parity batches would otherwise never take."""
    from oracle.molgym_oracle import CovariantOracle, ppo_loss
    from molgym_b200 import ppo
    monkeypatch.setenv('MGB_EDGE_MODE', mode)
    if mode == '0':   # together with the large-minibatch atom path (combined atom kernels, tiled InputLinear weight gradient)
        monkeypatch.setenv('MGB_SMALL_ATOMS', '0')
        monkeypatch.setenv('MGB_MIX_TC', '1')   # ... and the tensor-core channel mix (both directions)
        monkeypatch.setenv('MGB_LARGE_ATOMS', '1')   # ... and the rest of the large-minibatch choices (512-thread dcat mix, two-pass mix weight gradient)
    cfg = synth.CONFIGS['C3']
    torch.manual_seed(5)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    oracle.load_state_dict({k: v.detach().cpu() for k, v in agent.state_dict().items()})
    obs, n = synth.make_observations(cfg, batch=36)
    act = synth.make_actions(cfg, obs, n)
    ref = oracle.step(obs, act)
    old_logp, adv, ret = synth.make_ppo_targets(cfg, ref['logp'].detach().numpy())
    loss, _ = ppo.compute_loss(agent, dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret), 0.2, 0.5, 0.01)
    loss.backward()
    ref_loss, _ = ppo_loss(ref['logp'], ref['ent'], ref['v'], old_logp, adv, ret, 0.2, 0.5, 0.01)
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5
    ref_grads = {k: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k, p in oracle.named_parameters()}
    assert_grads_close(grads_of(agent), ref_grads)


def test_fused_ppo_step_matches_the_op_by_op_path():
    """ppo.compute_loss through CovariantAC.fused_ppo_loss (pinned staging copy + CUDA-graph replays of forward + k_ppo_loss and
    of the backward) against the same function with agent.fused_ppo = False (agent.step + torch ops + autograd): loss, info and
    every gradient; graphs are replayed on new canvases, after a parameter update, and with gradient accumulation."""
    from molgym_b200 import ppo
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    torch.manual_seed(3)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())

    def batch(seed, n_canvases=24):
        obs, n = synth.make_observations(cfg, batch=n_canvases, seed=seed)
        act = synth.make_actions(cfg, obs, n, seed=seed)
        with torch.no_grad():
            logp0 = agent.step(obs, act)['logp'].cpu().numpy()
        old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0, seed=seed)
        return dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)

    def run(data_list, fused):
        agent.fused_ppo = fused
        agent.zero_grad()
        out = []
        for data in data_list:
            loss, info = ppo.compute_loss(agent, data, 0.2, 0.5, 0.01)
            (loss * 0.5).backward()
            out.append((loss.item(), info))
        return out, {k: v.copy() for k, v in grads_of(agent).items()}

    for round_ in range(2):   # second round: same graphs replayed after the parameters moved
        data = [batch(100 + round_), batch(200 + round_)]
        ref_out, ref_grads = run(data, fused=False)
        got_out, got_grads = run(data, fused=True)
        for (l0, i0), (l1, i1) in zip(ref_out, got_out):
            assert abs(l0 - l1) <= 1e-6 * max(1.0, abs(l0))
            for key in i0:
                assert abs(i0[key] - i1[key]) <= 1e-6 * max(1.0, abs(i0[key])), key
        assert_grads_close(got_grads, ref_grads)
        with torch.no_grad():
            for p in agent.parameters():
                p.add_(0.01 * torch.randn_like(p))
    # consecutive fused steps alternate between two pipeline slots: two losses may be alive at once, a third call reuses the
    # first slot, and a loss superseded that way must raise instead of silently using the newer step's gradient
    agent.fused_ppo = True
    agent.zero_grad()
    d = batch(300)
    first, _ = ppo.compute_loss(agent, d, 0.2, 0.5, 0.01)
    second, _ = ppo.compute_loss(agent, d, 0.2, 0.5, 0.01)
    third, _ = ppo.compute_loss(agent, d, 0.2, 0.5, 0.01)
    with pytest.raises(RuntimeError):
        first.backward()
    second.backward()
    third.backward()
    twice = {k: v.copy() for k, v in grads_of(agent).items()}
    agent.fused_ppo = False
    agent.zero_grad()
    for _ in range(2):
        loss, _ = ppo.compute_loss(agent, d, 0.2, 0.5, 0.01)
        loss.backward()
    assert_grads_close(twice, grads_of(agent))
    # parameters edited on the caller's stream are seen by the next fused step (version counter of the flat buffer)
    agent.fused_ppo = True
    with torch.no_grad():
        for p in agent.parameters():
            p.mul_(1.01)
    agent.zero_grad()
    loss_f, _ = ppo.compute_loss(agent, d, 0.2, 0.5, 0.01)
    loss_f.backward()
    got = {k: v.copy() for k, v in grads_of(agent).items()}
    agent.fused_ppo = False
    agent.zero_grad()
    loss_u, _ = ppo.compute_loss(agent, d, 0.2, 0.5, 0.01)
    loss_u.backward()
    assert abs(loss_f.item() - loss_u.item()) <= 1e-6 * max(1.0, abs(loss_u.item()))
    assert_grads_close(got, grads_of(agent))


def test_gradients_accumulate_over_minibatches_like_autograd():
    """ppo.train sums gradients over minibatches before one optimizer step (ppo.py:118-131)."""
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    torch.manual_seed(2)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=40)
    act = synth.make_actions(cfg, obs, n)
    w = torch.linspace(-1, 1, 40, device='cuda')

    def objective(lo, hi):
        p = agent.step(obs[lo:hi], act[lo:hi])
        return (p['logp'] * w[lo:hi]).sum() + p['v'].sum() - 0.3 * p['ent'].sum()

    agent.zero_grad()
    objective(0, 40).backward()
    whole = grads_of(agent)
    agent.zero_grad()
    objective(0, 25).backward()
    objective(25, 40).backward()
    assert_grads_close(grads_of(agent), whole, rel=2e-5)
    # optimizer + clipping work on the parameters and update the flat buffer the kernels read
    opt = torch.optim.Adam(agent.parameters(), lr=1e-3)
    torch.nn.utils.clip_grad_norm_(agent.parameters(), 0.5)
    before = agent.step(obs, act)['logp'].detach().clone()
    opt.step()
    after = agent.step(obs, act)['logp'].detach()
    assert (before - after).abs().max() > 0


def test_full_size_properties_batch_independence_and_rotation_invariance():
    """C2 at the full minibatch (140): (1) evaluating canvases alone equals evaluating them in the batch; (2) a global
    rotation of canvas and orientation leaves logp / ent / v unchanged (SO(3) invariance, the property the reference's
    own agent tests pin, tests/agents/covariant/test_agent.py:43-123)."""
    cfg = synth.CONFIGS['C2']
    torch.manual_seed(4)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg)
    act = synth.make_actions(cfg, obs, n)
    with torch.no_grad():
        full = agent.step(obs, act)
        for lo in (0, 57, 133):
            part = agent.step(obs[lo:lo + 7], act[lo:lo + 7])
            for key in ('logp', 'ent', 'v'):
                assert torch.equal(part[key], full[key][lo:lo + 7]), key
        rng = np.random.default_rng(0)
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        rot_obs = [(tuple((lab, tuple(q @ np.asarray(xyz))) for lab, xyz in canvas), bag) for canvas, bag in obs]
        rot_act = act.copy()
        rot_act[:, 3:6] = act[:, 3:6] @ q.T
        rot = agent.step(rot_obs, rot_act)
    for key in ('logp', 'ent', 'v'):
        a, b = full[key].cpu().numpy(), rot[key].cpu().numpy()
        assert np.abs(a - b).max() <= 2e-4 * max(1.0, np.abs(a).max()), (key, np.abs(a - b).max())


def test_largest_canvas_config_runs_at_full_width_and_is_finite():
    cfg = synth.CONFIGS['C5']
    torch.manual_seed(1)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=256)
    act = synth.make_actions(cfg, obs, n)
    pred = agent.step(obs, act)
    (pred['logp'].mean() + pred['v'].mean()).backward()
    for key in ('logp', 'ent', 'v'):
        assert torch.isfinite(pred[key]).all()
    for name, p in agent.named_parameters():
        assert torch.isfinite(p.grad).all(), name


def test_rollout_mode_samples_valid_actions_and_consistent_logp():
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    torch.manual_seed(3)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=10)
    for training in (True, False):
        agent.training = training      # ppo.py:353,361 toggles the bare attribute
        with torch.no_grad():
            pred = agent.step(obs)
        a = pred['a'].cpu().numpy()
        assert a.shape == (10, 6) and len(pred['actions']) == 10
        for row, (canvas, bag), k in zip(a, obs, n):
            assert 0 <= round(row[0]) < max(k, 1)
            assert bag[int(round(row[1]))] > 0
            assert abs(np.linalg.norm(row[3:6]) - 1) < 1e-4
        with torch.no_grad():
            again = agent.step(obs, a)
        assert torch.allclose(again['logp'], pred['logp'], rtol=1e-5, atol=1e-5)


def test_whole_module_pickle_round_trip():
    """tools/model_util.py:82-117 saves and loads the whole module object."""
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=6)
    act = synth.make_actions(cfg, obs, n)
    clone = pickle.loads(pickle.dumps(agent))
    with torch.no_grad():
        assert torch.equal(agent.step(obs, act)['logp'], clone.step(obs, act)['logp'])


def test_bad_inputs_raise():
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=3)
    with pytest.raises(AssertionError):
        agent.step(obs, np.zeros((3, 5), np.float32))
    bad = [(((-1, (0.0, 0.0, 0.0)), ) + obs[0][0][1:], obs[0][1])]
    with pytest.raises(RuntimeError):
        agent.step(bad, np.zeros((1, 6), np.float32))


def test_bad_action_indices_raise_before_any_launch():
    """Discrete sub-actions index device arrays; the reference raises RuntimeError from to_one_hot (modules.py:8-23)."""
    from molgym_b200 import ppo
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=4)
    act = synth.make_actions(cfg, obs, n)
    for column, value in ((0, -1.0), (0, float(cfg.canvas_size)), (1, -1.0), (1, float(len(cfg.zs))), (0, float('nan'))):
        bad = act.copy()
        bad[2, column] = value
        with pytest.raises(RuntimeError):
            agent.step(obs, bad)
        with pytest.raises(RuntimeError):
            ppo.compute_loss(agent, dict(obs=obs, act=bad, logp=np.zeros(4, np.float32), adv=np.zeros(4), ret=np.zeros(4)), 0.2, 0.5, 0.01)


def test_unpickled_agent_sees_optimizer_updates_in_the_fused_step():
    """tools/model_util.py:93-117 reloads whole modules; the resumed agent's fused step (own streams) must be ordered behind the
    optimizer's in-place updates on the caller's stream exactly like a freshly built one."""
    from molgym_b200 import ppo
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    torch.manual_seed(6)
    agent = pickle.loads(pickle.dumps(make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())))
    assert agent._shared_version
    obs, n = synth.make_observations(cfg, batch=24)
    act = synth.make_actions(cfg, obs, n)
    with torch.no_grad():
        logp0 = agent.step(obs, act)['logp'].cpu().numpy()
    old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
    data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)
    opt = torch.optim.Adam(agent.parameters(), lr=1e-2)
    for _ in range(4):
        agent.fused_ppo = True
        agent.zero_grad()
        loss_f, info_f = ppo.compute_loss(agent, data, 0.2, 0.5, 0.01)
        loss_f.backward()
        fused = {k: v.copy() for k, v in grads_of(agent).items()}
        agent.fused_ppo = False
        agent.zero_grad()
        loss_u, info_u = ppo.compute_loss(agent, data, 0.2, 0.5, 0.01)
        loss_u.backward()
        assert abs(loss_f.item() - loss_u.item()) <= 1e-6 * max(1.0, abs(loss_u.item()))
        assert_grads_close(fused, grads_of(agent))
        version = agent._param_version()
        opt.step()
        assert agent._param_version() != version


def test_evaluate_slots_fall_back_when_busy_and_agree_with_eager_launches():
    """Evaluate-mode step() under autograd replays CUDA graphs on two persistent slots; a third live evaluation takes the eager
    path.  All three must give the same numbers and gradients as agent.graph_evaluate = False."""
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=64)
    torch.manual_seed(8)
    agent = make_agent(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=12)
    act = synth.make_actions(cfg, obs, n)
    w = torch.linspace(-1, 1, 12, device='cuda')

    def objective(p):
        return (p['logp'] * w).sum() + p['v'].sum() - 0.3 * p['ent'].sum()

    agent.graph_evaluate = False
    agent.zero_grad()
    ref = agent.step(obs, act)
    (3 * objective(ref)).backward()
    want = grads_of(agent)
    agent.graph_evaluate = True
    agent.zero_grad()
    preds = [agent.step(obs, act) for _ in range(3)]
    assert len(agent._eval_cache[12]) == 2
    for p in preds:
        for key in ('logp', 'ent', 'v'):
            assert torch.equal(p[key], ref[key]), key
    sum(objective(p) for p in preds).backward()
    assert_grads_close(grads_of(agent), want, rel=2e-5)
    # slots are free again: the next evaluations replay the graphs
    agent.zero_grad()
    objective(agent.step(obs, act)).backward()
    assert len(agent._eval_cache[12]) == 2


def test_c_abi_forward_backward_with_raw_device_pointers():
    """The contract a non-Python host binds (include/molgym_b200.h): plan, layout, workspace, forward, loss, backward called with
    raw device pointers through ctypes only, checked against the oracle."""
    import ctypes
    from molgym_b200 import _cabi, _lib
    from oracle.molgym_oracle import CovariantOracle, pack_observations, ppo_loss
    from scipy.integrate import lebedev_rule
    lib = _lib.load()
    cfg = dataclasses.replace(synth.CONFIGS['C3'], network_width=64)
    torch.manual_seed(21)
    oracle = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    c = _cabi.make_config(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    pts, wts = lebedev_rule(71)
    xyz, w = np.ascontiguousarray(pts.T, np.float64), np.ascontiguousarray(wts / (4 * np.pi), np.float64)
    plan = ctypes.c_void_p()
    dp = ctypes.POINTER(ctypes.c_double)
    _cabi.check(lib, lib.mgb_cov_plan_create(ctypes.byref(c), xyz.ctypes.data_as(dp), w.ctypes.data_as(dp), len(w), ctypes.byref(plan)))
    try:
        k = lib.mgb_cov_param_count(plan)
        off, num, tot = (ctypes.c_int64 * k)(), (ctypes.c_int64 * k)(), ctypes.c_int64()
        _cabi.check(lib, lib.mgb_cov_param_layout(plan, off, num, ctypes.byref(tot)))
        names = _cabi.param_names(cfg.num_cg_levels)
        state = oracle.state_dict()
        flat = np.zeros(tot.value, np.float32)
        for name, o, m in zip(names, off, num):
            flat[o:o + m] = state[name].numpy().ravel()
        B = 20
        obs, n = synth.make_observations(cfg, batch=B)
        act = synth.make_actions(cfg, obs, n)
        pos, charges, bags = pack_observations(obs, cfg.zs, cfg.canvas_size)
        ref = oracle.evaluate(pos, charges, bags, act)
        old_logp, adv, ret = synth.make_ppo_targets(cfg, ref['logp'].detach().numpy())
        dev = torch.device('cuda:0')
        d = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
        d_pos, d_ch, d_bags, d_act, d_flat = d(pos.astype(np.float32)), d(charges.astype(np.int32)), d(bags.astype(np.float32)), d(act), d(flat)
        d_old, d_adv, d_ret = d(old_logp), d(adv), d(ret)
        out = torch.zeros(6, B, device=dev)
        info = torch.zeros(8, dtype=torch.float64, device=dev)
        grad = torch.full((tot.value, ), 7.0, device=dev)     # accumulate = 0 must overwrite
        nbytes = lib.mgb_cov_workspace_bytes(plan, B)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        o = _cabi.CovOutputs()
        o.logp, o.ent, o.v = out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr()
        stream = torch.cuda.current_stream().cuda_stream
        assert lib.mgb_cov_forward(plan, B, d_pos.data_ptr(), d_ch.data_ptr(), d_bags.data_ptr(), d_act.data_ptr(), d_flat.data_ptr(),
                                   ws.data_ptr(), nbytes - 1, ctypes.byref(o), stream) == -3       # MGB_ERR_WORKSPACE
        assert b'workspace' in lib.mgb_last_error()
        _cabi.check(lib, lib.mgb_cov_forward(plan, B, d_pos.data_ptr(), d_ch.data_ptr(), d_bags.data_ptr(), d_act.data_ptr(),
                                             d_flat.data_ptr(), ws.data_ptr(), nbytes, ctypes.byref(o), stream))
        _cabi.check(lib, lib.mgb_ppo_loss(B, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), d_old.data_ptr(), d_adv.data_ptr(),
                                          d_ret.data_ptr(), 0.2, 0.5, 0.01, 1.0 / B, info.data_ptr(), out[3].data_ptr(), out[4].data_ptr(),
                                          out[5].data_ptr(), stream))
        _cabi.check(lib, lib.mgb_cov_backward(plan, B, d_pos.data_ptr(), d_ch.data_ptr(), d_bags.data_ptr(), d_act.data_ptr(),
                                              d_flat.data_ptr(), ws.data_ptr(), nbytes, out[3].data_ptr(), out[4].data_ptr(),
                                              out[5].data_ptr(), grad.data_ptr(), 0, stream))
        torch.cuda.synchronize()
        for idx, key in enumerate(('logp', 'ent', 'v')):
            assert_outputs_close(out[idx].cpu().numpy(), ref[key].detach().numpy(), rel=OUT_REL, what=key)
        ref_loss, _ = ppo_loss(ref['logp'], ref['ent'], ref['v'], old_logp, adv, ret, 0.2, 0.5, 0.01)
        ref_loss.backward()
        assert abs(float(info[0]) - ref_loss.item()) <= 1e-5 * max(1.0, abs(ref_loss.item()))
        g = grad.cpu().numpy()
        got = {name: g[o_:o_ + m].reshape(tuple(state[name].shape)) for name, o_, m in zip(names, off, num)}
        ref_grads = {k_: (p.grad.numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32)) for k_, p in oracle.named_parameters()}
        assert_grads_close(got, ref_grads)
    finally:
        lib.mgb_cov_plan_destroy(plan)
