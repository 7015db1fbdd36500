"""TEST INFRASTRUCTURE — the product's agent classes on top of the kernel emulator build.

`EmuCovariantAC` / `EmuSchNetAC` are the unmodified molgym_b200 agents with one substitution: their runtime object
(molgym_b200/agents/_runtime.py) is replaced by `HostRuntime` — CPU tensors, the g++ -DMGB_CUSIM build of the same .cu
sources (tests/cusim/_build), streams / events that do nothing and "graphs" that simply re-issue their launch sequence.
Everything above the C ABI (packing, sharding, slot bookkeeping, gradient accumulation, pickling, the autograd glue) is the
product's own code, which lets the CPU-only container drive the reference's unchanged ppo.train on it and run world-size-2
gloo tests.  The product never imports this module."""
import contextlib

import torch

from molgym_b200.agents.covariant.agent import CovariantAC
from molgym_b200.agents.internal.agent import SchNetAC
from tests.cusim import runner


class _HostStream:
    cuda_stream = None

    def wait_stream(self, other):
        pass

    def wait_event(self, event):
        pass

    def synchronize(self):
        pass


class _HostEvent:
    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


class _HostGraph:
    def __init__(self, fn):
        self.fn = fn
        fn(None)

    def replay(self):
        self.fn(None)


class HostRuntime:
    is_cuda = False

    def __init__(self, device=None):
        self.device = torch.device('cpu')
        self._stream = _HostStream()

    def lib(self):
        return runner.lib()

    def device_ctx(self):
        return contextlib.nullcontext()

    def current_stream(self):
        return self._stream

    def stream_ptr(self, stream=None):
        return None

    def new_stream(self, priority=0):
        return _HostStream()

    def new_event(self):
        return _HostEvent()

    def stream_ctx(self, stream):
        return contextlib.nullcontext()

    def pinned(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8)

    def capture(self, fn):
        return _HostGraph(fn)

    def total_memory(self):
        return 1 << 40


class EmuCovariantAC(CovariantAC):
    _runtime_cls = HostRuntime


class EmuSchNetAC(SchNetAC):
    _runtime_cls = HostRuntime


def make_emu_case(width=32):
    """(cfg, agent, data): a small covariant agent on the emulator and one synthetic PPO minibatch of 11 canvases."""
    import dataclasses

    from molgym_b200 import synth
    from molgym_b200.spaces import ActionSpace, ObservationSpace
    cfg = dataclasses.replace(synth.CONFIGS['C2'], network_width=width, canvas_size=5, bag={16: 1, 9: 4})
    torch.manual_seed(0)
    agent = EmuCovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=11)
    act = synth.make_actions(cfg, obs, n)
    with torch.no_grad():
        logp0 = agent.step(obs, act)['logp'].numpy()
    old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
    return cfg, agent, dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)


def run_ppo_epoch(agent, data, fused, splits=((0, 6), (6, 11))):
    """ppo.train's inner loop (ppo.py:118-135) on two minibatches: zero_grad, compute_loss + backward per minibatch (gradients
    accumulate), then the gradients are read through agent.parameters()."""
    from molgym_b200 import ppo
    agent.fused_ppo = fused
    agent.zero_grad()
    infos = []
    for lo, hi in splits:
        batch = {k: v[lo:hi] for k, v in data.items()}
        loss, info = ppo.compute_loss(agent, batch, 0.2, 0.5, 0.01)
        loss.backward()
        infos.append((float(loss.item()), dict(info)))
    grads = torch.cat([p.grad.reshape(-1) for p in agent.parameters()]).clone()
    return infos, grads


def run_train(agent, data, steps=2, mini_batch_size=6):
    """molgym_b200.ppo.train for `steps` optimizer steps (plain SGD, so that the parameter trajectory is easy to compare):
    returned infos without the wall-clock entry, final flat parameters."""
    import numpy as np

    from molgym_b200 import ppo
    agent.fused_ppo = True
    np.random.seed(7)   # get_batch_generator permutes with numpy's global generator
    opt = torch.optim.SGD(torch.nn.Module.parameters(agent), lr=0.05)
    infos = ppo.train(agent, opt, data, mini_batch_size=mini_batch_size, clip_ratio=0.2, target_kl=1e9, vf_coef=0.5, entropy_coef=0.01,
                      gradient_clip=0.5, max_num_steps=steps)
    infos.pop('time', None)
    return {k: float(v) for k, v in infos.items()}, agent._flat.detach().numpy().copy()


def sharded_worker(rank, world, port, out):
    """World-size-2 gloo worker: the data-parallel agent on every rank is handed the SAME minibatches (as under torchrun with
    the unchanged ppo.train) and evaluates its shard."""
    import os

    import torch.distributed as dist

    from molgym_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg, agent, data = make_emu_case()
    if rank == 1:   # a rank that starts from different parameters is brought in line by shard_agent's broadcast
        with torch.no_grad():
            for p in torch.nn.Module.parameters(agent):
                p.add_(0.01)
    parallel.shard_agent(agent)
    res = {}
    for fused in (True, False):
        infos, grads = run_ppo_epoch(agent, data, fused)
        res[fused] = (infos, grads.numpy())
    # a pending reduction is also flushed by an optimizer step that never looked at parameters() first
    opt = torch.optim.SGD(torch.nn.Module.parameters(agent), lr=0.1)
    agent.zero_grad()
    from molgym_b200 import ppo
    agent.fused_ppo = True
    loss, _ = ppo.compute_loss(agent, data, 0.2, 0.5, 0.01)
    loss.backward()
    pending_before = agent._grad_pending
    opt.step()
    res['hook'] = (pending_before, agent._grad_pending, agent._flat.detach().numpy().copy())
    # the restated ppo.train on the sharded agent: every rank collects only its slice of each minibatch (train_shard / n_global)
    _, agent_t, data_t = make_emu_case()
    parallel.shard_agent(agent_t)
    res['train'] = run_train(agent_t, data_t)
    out[rank] = res
    dist.destroy_process_group()
