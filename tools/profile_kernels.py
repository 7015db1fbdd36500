"""Warm, in-situ per-kernel times of the device-resident step: CUDA events around every launch (mgb_profile_kernel('k_')),
on the stream each kernel is launched on, L2 flushed before each step like bench.py.  Unlike an ncu launch list the kernels
run back to back with the producers' outputs still in L2 and the side streams overlapping.

    python tools/profile_kernels.py C2 20 [batch]
"""
import collections
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molgym_b200 import _cabi, _lib, synth  # noqa: E402
from molgym_b200.agents.covariant.agent import CovariantAC  # noqa: E402
from molgym_b200.spaces import ActionSpace, ObservationSpace  # noqa: E402


def main(workload='C2', steps=20, batch=None):
    steps = int(steps)
    cfg = synth.CONFIGS[workload]
    B = int(batch) if batch else cfg.mini_batch_size
    dev = torch.device('cuda:0')
    lib = _lib.load()
    torch.manual_seed(0)
    agent = CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=B)
    act = synth.make_actions(cfg, obs, n)
    parsed = agent.parse_observations(obs)
    pos, charges, bags = parsed['positions'], parsed['charges'], parsed['bags']
    act_d = torch.as_tensor(act, dtype=torch.float32, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    logp, ent, v = torch.empty(B, **f32), torch.empty(B, **f32), torch.empty(B, **f32)
    g = [torch.empty(B, **f32) for _ in range(3)]
    old = torch.zeros(B, **f32)
    adv = torch.randn(B, dtype=torch.float64, device=dev)
    ret = torch.randn(B, dtype=torch.float64, device=dev)
    info = torch.zeros(8, dtype=torch.float64, device=dev)
    grad = torch.zeros_like(agent._flat)
    ws = torch.empty(lib.mgb_cov_workspace_bytes(agent._plan, B), dtype=torch.uint8, device=dev)
    outs = _cabi.CovOutputs()
    outs.logp, outs.ent, outs.v = logp.data_ptr(), ent.data_ptr(), v.data_ptr()
    stream = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        _cabi.check(lib, lib.mgb_cov_forward(agent._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(), act_d.data_ptr(),
                                             agent._flat.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(outs), stream))
        _cabi.check(lib, lib.mgb_ppo_loss(B, logp.data_ptr(), ent.data_ptr(), v.data_ptr(), old.data_ptr(), adv.data_ptr(), ret.data_ptr(),
                                          0.2, 0.5, 0.01, 1.0 / B, info.data_ptr(), g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), stream))
        _cabi.check(lib, lib.mgb_cov_backward(agent._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(), act_d.data_ptr(),
                                              agent._flat.data_ptr(), ws.data_ptr(), ws.numel(), g[0].data_ptr(), g[1].data_ptr(),
                                              g[2].data_ptr(), grad.data_ptr(), 0, stream))

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    tot = collections.OrderedDict()
    buf = ctypes.create_string_buffer(1 << 20)
    t_step = 0.0
    for s in range(steps):
        flush.zero_()
        lib.mgb_profile_kernel(b'k_')
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        t_step += a.elapsed_time(b)
        lib.mgb_profile_report(buf, len(buf))
        seen = collections.Counter()
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit(' ', 1)
            seen[name] += 1
            key = '%s #%d' % (name, seen[name])
            tot[key] = tot.get(key, 0.0) + float(ms)
    lib.mgb_profile_kernel(None)
    ssum = sum(tot.values()) / steps
    print('# %s B=%d: step (events around every launch) %.1f us; sum of kernel times %.1f us over %d launches'
          % (workload, B, t_step / steps * 1e3, ssum * 1e3, len(tot)))
    for k, v_ in tot.items():
        print('%-44s %9.1f us %5.1f%%' % (k, v_ / steps * 1e3, 100 * v_ / steps / ssum))


if __name__ == '__main__':
    main(*sys.argv[1:])
