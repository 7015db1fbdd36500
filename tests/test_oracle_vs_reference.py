"""Build-container-only checks (skipped where /root/reference is absent, e.g. on the GPU box):
 1. the reference's OWN hot-path tests pass, unmodified, on the restated third-party stand-ins;
 2. the self-contained oracle equals the reference's own code run verbatim on fresh canvases."""
import dataclasses
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import refrun

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_reference = pytest.mark.skipif(not refrun.reference_available(), reason='/root/reference not present')


@needs_reference
def test_reference_own_tests_pass_on_standins():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'oracle', 'run_reference_tests.py')],
                          capture_output=True, text=True, timeout=900)
    tail = proc.stdout[-2000:]
    assert proc.returncode == 0, tail
    assert ' passed' in tail and 'failed' not in tail


@needs_reference
@pytest.mark.parametrize('beta', [-10.0, None])
def test_oracle_equals_verbatim_reference(beta):
    refrun.enable(require_reference=True)
    from molgym.agents.covariant.agent import CovariantAC
    from molgym.spaces import ActionSpace, ObservationSpace
    from molgym.tools import util

    from molgym_b200 import synth
    from oracle.molgym_oracle import CovariantOracle

    cfg = dataclasses.replace(synth.CONFIGS['C3'], beta=beta, network_width=32, seed=5)
    util.set_seeds(3)
    osp = ObservationSpace(canvas_size=cfg.canvas_size, zs=cfg.zs)
    ref = CovariantAC(observation_space=osp, action_space=ActionSpace(zs=cfg.zs), device=torch.device('cpu'),
                      **cfg.agent_kwargs())
    orc = CovariantOracle(cfg.zs, cfg.canvas_size, **cfg.agent_kwargs())
    orc.load_state_dict(ref.state_dict())
    obs, n = synth.make_observations(cfg, batch=13)
    act = synth.make_actions(cfg, obs, n)
    r = ref.step(obs, act)
    o = orc.step(obs, act)
    for k in ('logp', 'ent', 'v'):
        np.testing.assert_allclose(o[k].detach().numpy(), r[k].detach().numpy(), rtol=2e-6, atol=2e-6)
