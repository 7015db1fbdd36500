"""The self-contained oracle must reproduce the golden vectors that the reference's own code produced
(tests/golden/make_golden.py).  CPU only; this is what pins the oracle on the GPU box."""
import numpy as np
import pytest
import torch

from oracle.molgym_oracle import CovariantOracle, ppo_loss
from tests.util_golden import (agent_kwargs_from_config, assert_outputs_close, golden_grads, golden_observations,
                               golden_state_dict, load_golden, assert_grads_close)

CASES = ['covariant_sf6_beta', 'covariant_hco_nobeta_trained']


@pytest.mark.parametrize('name', CASES)
def test_oracle_reproduces_reference_golden(name):
    g = load_golden(name)
    cfg = g['config']
    oracle = CovariantOracle(cfg['zs'], cfg['canvas_size'], **agent_kwargs_from_config(cfg))
    missing = oracle.load_state_dict(golden_state_dict(g))
    assert not missing.missing_keys and not missing.unexpected_keys
    out = oracle.step(golden_observations(g), g['actions'])
    for key in ('logp', 'ent', 'v', 'focus_probs', 'element_probs'):
        assert_outputs_close(out[key].detach().numpy(), g[key], rel=2e-6, what=key)
    for ell, part in enumerate(out['coefficients']):
        assert_outputs_close(part.detach().numpy(), g[f'coeff_{ell}'], rel=2e-6, what=f'coeff_{ell}')
    if 'log_z' in g:
        assert_outputs_close(out['log_z'].detach().numpy(), g['log_z'], rel=2e-6, what='log_z')
    loss, info = ppo_loss(out['logp'], out['ent'], out['v'], g['old_logp'], g['adv'], g['ret'], 0.2, 0.5, 0.01)
    assert abs(loss.item() - float(g['loss'])) <= 1e-6 * max(1.0, abs(float(g['loss'])))
    for key in ('policy_loss', 'vf_loss', 'entropy_loss', 'approx_kl', 'clip_fraction'):
        assert abs(info[key] - float(g['info_' + key])) <= 1e-6
    loss.backward()
    got = {n: (p.grad.numpy() if p.grad is not None else np.zeros(p.shape, np.float32))
           for n, p in oracle.named_parameters()}
    assert_grads_close(got, golden_grads(g))


def test_fp64_oracle_is_close_to_fp32_golden():
    """The float64 arbiter agrees with the float32 reference to float32 round-off."""
    g = load_golden('covariant_sf6_beta')
    cfg = g['config']
    oracle = CovariantOracle(cfg['zs'], cfg['canvas_size'], dtype=torch.float64, **agent_kwargs_from_config(cfg))
    oracle.load_state_dict({k: v.double() for k, v in golden_state_dict(g).items()})
    out = oracle.step(golden_observations(g), g['actions'])
    for key in ('logp', 'ent', 'v'):
        assert_outputs_close(out[key].detach().numpy(), g[key], rel=2e-5, floor=5e-2, what=key)
