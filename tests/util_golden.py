"""Helpers shared by the parity tests: load golden vectors, rebuild observations, load parameters."""
import ast
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g['config'] = ast.literal_eval(str(g.pop('config_json')))
    return g


def golden_observations(g):
    obs = []
    for labels, xyz, bag in zip(g['labels'], g['xyz'], g['bags']):
        canvas = tuple((int(lab), tuple(float(x) for x in p)) for lab, p in zip(labels, xyz))
        obs.append((canvas, tuple(int(c) for c in bag)))
    return obs


def golden_state_dict(g):
    return {k[len('param/'):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith('param/')}


def golden_grads(g):
    return {k[len('grad/'):]: v for k, v in g.items() if k.startswith('grad/')}


def agent_kwargs_from_config(cfg):
    keys = ('min_max_distance', 'network_width', 'maxl', 'num_cg_levels', 'num_channels_hidden',
            'num_channels_per_element', 'num_gaussians', 'bag_scale', 'beta')
    kw = {k: cfg[k] for k in keys}
    kw['min_max_distance'] = tuple(kw['min_max_distance'])
    return kw


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    denom = max(np.linalg.norm(b), 1e-30)
    return float(np.linalg.norm(a - b) / denom)


def assert_outputs_close(x, ref, rel=1e-5, floor=1e-3, what=''):
    """abs(x - ref) <= rel * max(abs(ref), floor)  (SURVEY.md section 8c parity protocol)."""
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    tol = rel * np.maximum(np.abs(ref), floor)
    bad = np.abs(x - ref) > tol
    assert not bad.any(), f'{what}: max abs err {np.abs(x - ref).max():.3e}, worst tol {tol.min():.3e}, ' \
                          f'{bad.sum()} / {bad.size} outside tolerance'


def assert_grads_close(got: dict, ref: dict, rel=1e-4, floor_frac=1e-3):
    """Per-tensor ||got - ref|| <= rel * max(||ref||, floor_frac * G), G = largest per-tensor ||ref||.
    (Some gradients are mathematically zero — e.g. the last focus bias, by softmax shift invariance — and
    only carry round-off noise, hence the floor.)"""
    G = max(float(np.linalg.norm(np.asarray(v, dtype=np.float64))) for v in ref.values())
    worst = (0.0, None)
    for name, r in ref.items():
        r = np.asarray(r, dtype=np.float64)
        x = np.asarray(got[name], dtype=np.float64)
        err = float(np.linalg.norm(x - r))
        tol = rel * max(float(np.linalg.norm(r)), floor_frac * G)
        if err / tol > worst[0]:
            worst = (err / tol, name)
        assert err <= tol, f'grad {name}: ||err||={err:.3e} > tol={tol:.3e} (||ref||={np.linalg.norm(r):.3e}, G={G:.3e})'
    return worst
