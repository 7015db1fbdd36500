"""TEST INFRASTRUCTURE — drive the C ABI of the kernel emulator build (tests/cusim/_build/) with numpy arrays.

The same .cu sources as the product, compiled by g++ against cusim.h: lets the CPU-only container check kernel logic
(indexing, barriers, reductions, hand-written backward) against the oracle before GPU time is spent.
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from molgym_b200 import _cabi, build  # noqa: E402

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = build.build_cusim()
        _LIB = _cabi.bind(ctypes.CDLL(path))
        assert _LIB.mgb_is_cuda_build() == 0
    return _LIB


def lebedev():
    from scipy.integrate import lebedev_rule
    pts, w = lebedev_rule(71)
    return np.ascontiguousarray(pts.T, dtype=np.float64), np.ascontiguousarray(w / (4 * np.pi), dtype=np.float64)


def ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class CusimCov:
    def __init__(self, zs, canvas_size, **kw):
        self.L = lib()
        self.cfg = _cabi.make_config(zs, canvas_size, **kw)
        xyz, w = lebedev()
        plan = ctypes.c_void_p()
        _cabi.check(self.L, self.L.mgb_cov_plan_create(ctypes.byref(self.cfg), xyz.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                       w.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(w),
                                                       ctypes.byref(plan)))
        self.plan = plan
        n = self.L.mgb_cov_param_count(plan)
        off = (ctypes.c_int64 * n)()
        num = (ctypes.c_int64 * n)()
        tot = ctypes.c_int64()
        _cabi.check(self.L, self.L.mgb_cov_param_layout(plan, off, num, ctypes.byref(tot)))
        self.offsets, self.numels, self.total = list(off), list(num), tot.value
        self.names = _cabi.param_names(self.cfg.num_cg_levels)
        assert len(self.names) == n, (len(self.names), n)
        self.N, self.Z = canvas_size, len(zs)
        self.CPE, self.G = self.cfg.num_channels_per_element, self.cfg.num_gaussians
        self.ws = None

    def __del__(self):
        try:
            self.L.mgb_cov_plan_destroy(self.plan)
        except Exception:
            pass

    def flatten(self, state_dict):
        flat = np.zeros(self.total, dtype=np.float32)
        for name, o, n in zip(self.names, self.offsets, self.numels):
            v = np.asarray(state_dict[name].detach().cpu().numpy() if hasattr(state_dict[name], 'detach') else state_dict[name],
                           dtype=np.float32).ravel()
            assert v.size == n, (name, v.size, n)
            flat[o:o + n] = v
        return flat

    def unflatten(self, flat, shapes):
        return {name: flat[o:o + n].reshape(shapes[name]) for name, o, n in zip(self.names, self.offsets, self.numels)}

    def forward(self, pos, charges, bags, actions, params):
        B = len(pos)
        self.B = B
        self.inputs = [np.ascontiguousarray(pos, np.float32), np.ascontiguousarray(charges, np.int32),
                       np.ascontiguousarray(bags, np.float32), np.ascontiguousarray(actions, np.float32),
                       np.ascontiguousarray(params, np.float32)]
        nbytes = self.L.mgb_cov_workspace_bytes(self.plan, B)
        self.ws = np.zeros(nbytes // 4 + 64, dtype=np.float32)
        self.ws_bytes = nbytes
        N, Z = self.N, self.Z
        o = dict(logp=np.zeros(B, np.float32), ent=np.zeros(B, np.float32), v=np.zeros(B, np.float32),
                 logp_parts=np.zeros((B, 4), np.float32), focus_probs=np.zeros((B, N), np.float32),
                 element_probs=np.zeros((B, Z), np.float32), gmm=np.zeros((B, 3, self.G), np.float32),
                 coefficients=np.zeros((B, 25, self.CPE, 2), np.float32), log_z=np.zeros(B, np.float32),
                 covariats=np.zeros((B, N, 25, Z * self.CPE, 2), np.float32))
        outs = _cabi.CovOutputs(**{k: ptr(v) for k, v in o.items()})
        _cabi.check(self.L, self.L.mgb_cov_forward(self.plan, B, *[ptr(a) for a in self.inputs], ptr(self.ws), nbytes,
                                                   ctypes.byref(outs), None))
        return o

    def backward(self, g_logp, g_ent, g_v, accumulate_into=None):
        grad = np.zeros(self.total, np.float32) if accumulate_into is None else accumulate_into
        g = [np.ascontiguousarray(x, np.float32) for x in (g_logp, g_ent, g_v)]
        _cabi.check(self.L, self.L.mgb_cov_backward(self.plan, self.B, *[ptr(a) for a in self.inputs], ptr(self.ws),
                                                    self.ws_bytes, ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(grad),
                                                    0 if accumulate_into is None else 1, None))
        return grad


def ppo_loss(logp, ent, v, old_logp, adv, ret, clip, vf, ent_coef, inv_b=None):
    L = lib()
    B = len(logp)
    a = [np.ascontiguousarray(x, np.float32) for x in (logp, ent, v, old_logp)]
    d = [np.ascontiguousarray(x, np.float64) for x in (adv, ret)]
    info = np.zeros(8, np.float64)
    g = [np.zeros(B, np.float32) for _ in range(3)]
    _cabi.check(L, L.mgb_ppo_loss(B, *[ptr(x) for x in a], *[ptr(x) for x in d], clip, vf, ent_coef, (1.0 / B) if inv_b is None else inv_b,
                                  ptr(info), ptr(g[0]), ptr(g[1]), ptr(g[2]), None))
    return info, g


def scale_accumulate(dst, src, scale, accumulate):
    """mgb_scale_accumulate on host arrays (dst / src must be 16-byte aligned: numpy allocations are)."""
    L = lib()
    dst = np.ascontiguousarray(dst, np.float32).copy()
    src = np.ascontiguousarray(src, np.float32)
    sc = np.asarray([scale])
    assert sc.dtype in (np.float32, np.float64)
    _cabi.check(L, L.mgb_scale_accumulate(ptr(dst), ptr(src), ptr(sc), 1 if sc.dtype == np.float64 else 0, dst.size, int(accumulate), None))
    return dst


def pack(zs, canvas_size, labels, xyz):
    L = lib()
    cfg = _cabi.CovConfig()
    cfg.canvas_size, cfg.num_species = canvas_size, len(zs)
    for i, z in enumerate(zs):
        cfg.zs[i] = int(z)
    labels = np.ascontiguousarray(labels, np.int32)
    xyz = np.ascontiguousarray(xyz, np.float64)
    B = len(labels)
    pos = np.empty((B, canvas_size, 3), np.float32)
    charges = np.empty((B, canvas_size), np.int32)
    rc = L.mgb_pack_observations(ctypes.byref(cfg), B, ptr(labels), ptr(xyz), ptr(pos), ptr(charges))
    if rc != 0:
        raise RuntimeError(L.mgb_last_error().decode())
    return pos, charges


class CusimInt:
    """Internal-coordinate (SchNet) agent through the emulator build."""

    def __init__(self, zs, canvas_size, min_max_distance, network_width):
        self.L = lib()
        self.cfg = _cabi.make_int_config(zs, canvas_size, min_max_distance, network_width)
        plan = ctypes.c_void_p()
        _cabi.check(self.L, self.L.mgb_int_plan_create(ctypes.byref(self.cfg), ctypes.byref(plan)))
        self.plan = plan
        n = self.L.mgb_int_param_count(plan)
        off = (ctypes.c_int64 * n)()
        num = (ctypes.c_int64 * n)()
        tot = ctypes.c_int64()
        _cabi.check(self.L, self.L.mgb_int_param_layout(plan, off, num, ctypes.byref(tot)))
        self.offsets, self.numels, self.total = list(off), list(num), tot.value
        self.names = _cabi.int_param_names()
        assert len(self.names) == n, (len(self.names), n)
        self.N, self.Z = canvas_size, len(zs)

    flatten = CusimCov.flatten
    unflatten = CusimCov.unflatten

    def forward(self, numbers, positions, bags, actions, params):
        B = len(numbers)
        self.B = B
        self.inputs = [np.ascontiguousarray(numbers, np.int32), np.ascontiguousarray(positions, np.float32),
                       np.ascontiguousarray(bags, np.float32), np.ascontiguousarray(actions, np.float32),
                       np.ascontiguousarray(params, np.float32)]
        nbytes = self.L.mgb_int_workspace_bytes(self.plan, B)
        self.ws = np.zeros(nbytes // 4 + 64, dtype=np.float32)
        self.ws_bytes = nbytes
        o = dict(logp=np.zeros(B, np.float32), ent=np.zeros(B, np.float32), v=np.zeros(B, np.float32),
                 logp_terms=np.zeros((B, 6), np.float32), focus_probs=np.zeros((B, self.N), np.float32),
                 element_probs=np.zeros((B, self.Z), np.float32), means=np.zeros((B, 3), np.float32),
                 kappa_logits=np.zeros((B, 2), np.float32))
        outs = _cabi.IntOutputs(**{k: ptr(v) for k, v in o.items()})
        _cabi.check(self.L, self.L.mgb_int_forward(self.plan, B, *[ptr(a) for a in self.inputs], ptr(self.ws), nbytes, ctypes.byref(outs), None))
        return o

    def backward(self, g_logp, g_ent, g_v):
        grad = np.zeros(self.total, np.float32)
        g = [np.ascontiguousarray(x, np.float32) for x in (g_logp, g_ent, g_v)]
        _cabi.check(self.L, self.L.mgb_int_backward(self.plan, self.B, *[ptr(a) for a in self.inputs], ptr(self.ws), self.ws_bytes,
                                                    ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(grad), 0, None))
        return grad
