"""Data-parallel sharding of the PPO minibatch: one process per GPU, every rank runs the same
(unchanged) ppo.train loop on the same rollout buffer, and a data-parallel agent (`shard_agent`) evaluates only its
contiguous shard of each minibatch:

  * forward / backward: canvases are independent, no data-path collective;
  * loss: each rank's PPO terms are sums over its shard divided by the GLOBAL minibatch size; one all-reduce of the 8-double
    info block (fused step) or an autograd-aware all-gather of logp / ent / v (evaluate-mode step(), 12 bytes per canvas)
    gives every rank the global loss, approx_kl and clip_fraction, so all ranks take the same early-stop branch
    (molgym/ppo.py:138-140);
  * gradient: the shard's gradient accumulates locally over the minibatches of an epoch and is summed over ranks ONCE per
    optimizer step (agents/_flat.py::sync_grads, NCCL over NVLink) — the path's one exchange step."""
from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n canvases into `world` shards whose sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_agent(agent, broadcast: bool = True):
    """Mark the agent data-parallel and make every rank start from rank 0's parameters."""
    agent.data_parallel = True
    if broadcast and dist.is_initialized() and dist.get_world_size() > 1:
        agent._realias()
        dist.broadcast(agent._flat, src=0)
    return agent


class _GatherShards(torch.autograd.Function):
    """all-gather of per-shard [n_r] vectors into the global [n] vector; backward hands each rank its slice."""

    @staticmethod
    def forward(ctx, local: torch.Tensor, n: int):
        world, rank = dist.get_world_size(), dist.get_rank()
        ctx.bounds = shard_bounds(n, rank, world)
        width = (n + world - 1) // world
        padded = local.new_zeros(width)
        padded[:local.numel()] = local
        gathered = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(gathered, padded)
        parts = []
        for r in range(world):
            lo, hi = shard_bounds(n, r, world)
            parts.append(gathered[r][:hi - lo])
        return torch.cat(parts)

    @staticmethod
    def backward(ctx, grad):
        lo, hi = ctx.bounds
        return grad[lo:hi].contiguous(), None


def gather_shards(local: torch.Tensor, n: int) -> torch.Tensor:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    return _GatherShards.apply(local, n)


def global_step(agent, observations: List, actions) -> dict:
    """step() on the whole minibatch with the work sharded over ranks; returns global logp / ent / v.  (A data-parallel
    agent's own step() does exactly this; kept for callers that shard an agent they did not mark data-parallel.)"""
    if getattr(agent, 'data_parallel', False):
        return agent.step(observations, actions)
    n = len(observations)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(n, rank, world)
    pred = agent.step(observations[lo:hi], actions[lo:hi])
    out = dict(pred)
    for k in ('logp', 'ent', 'v'):
        out[k] = gather_shards(pred[k], n)
    return out


def _gloo_selftest_worker(rank, world, port, n, out):
    """Worker of tests/test_parallel_gloo.py (lives here so that spawned processes can import it by an unambiguous module
    path): a stand-in "logp" of the local shard is gathered, a loss over the GLOBAL batch is differentiated, and the
    parameter gradients are summed over ranks — the data-parallel path's host logic, on CPU with gloo."""
    import os
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    theta = torch.tensor([0.3, -0.7], requires_grad=True)
    x = torch.linspace(-1, 1, n)
    lo, hi = shard_bounds(n, rank, world)
    local = torch.sin(theta[0] * x[lo:hi]) + theta[1] * x[lo:hi]**2
    full = gather_shards(local, n)
    loss = (full * torch.cos(x)).mean()
    loss.backward()
    g = theta.grad.clone()
    dist.all_reduce(g)
    out[rank] = (loss.item(), g.numpy().tolist(), full.detach().numpy().tolist())
    dist.destroy_process_group()
