"""Run only the device-resident hot path (mgb_cov_forward -> mgb_ppo_loss -> mgb_cov_backward) a few times, for ncu:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_step.py C2 3
Prints the number of kernel launches per step so that -s / -c can select whole steps."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from molgym_b200 import _cabi, _lib, synth  # noqa: E402
from molgym_b200.agents.covariant.agent import CovariantAC  # noqa: E402
from molgym_b200.spaces import ActionSpace, ObservationSpace  # noqa: E402


def main(workload='C2', steps=3, batch=None):
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
    torch.cuda.synchronize()
    n0 = lib.mgb_launch_count()
    for s in range(int(steps)):
        _cabi.check(lib, lib.mgb_cov_forward(agent._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(), act_d.data_ptr(),
                                             agent._flat.data_ptr(), ws.data_ptr(), ws.numel(), ctypes.byref(outs), stream))
        _cabi.check(lib, lib.mgb_ppo_loss(B, logp.data_ptr(), ent.data_ptr(), v.data_ptr(), old.data_ptr(), adv.data_ptr(), ret.data_ptr(),
                                          0.2, 0.5, 0.01, 1.0 / B, info.data_ptr(), g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), stream))
        _cabi.check(lib, lib.mgb_cov_backward(agent._plan, B, pos.data_ptr(), charges.data_ptr(), bags.data_ptr(), act_d.data_ptr(),
                                              agent._flat.data_ptr(), ws.data_ptr(), ws.numel(), g[0].data_ptr(), g[1].data_ptr(),
                                              g[2].data_ptr(), grad.data_ptr(), 0, stream))
        torch.cuda.synchronize()
        if s == 0:
            print('kernel launches per step:', lib.mgb_launch_count() - n0, flush=True)


if __name__ == '__main__':
    main(*sys.argv[1:])
