"""Hardware data-parallel parity (needs >= 2 CUDA devices: run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`; skipped on
the single-GPU test box): a PPO minibatch of 64 canvases sharded 2 x 32 over two NCCL ranks through the fused CUDA-graph step and
through evaluate-mode step() must give every rank bit-identical loss info and, after the one deferred all-reduce, the single-GPU
gradient."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    import dataclasses

    import torch.distributed as dist

    from molgym_b200 import parallel, ppo, synth
    from molgym_b200.agents.covariant.agent import CovariantAC
    from molgym_b200.spaces import ActionSpace, ObservationSpace
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    cfg = dataclasses.replace(synth.CONFIGS['C3'], network_width=64)
    torch.manual_seed(0)
    agent = CovariantAC(ObservationSpace(cfg.canvas_size, cfg.zs), ActionSpace(cfg.zs), device=dev, **cfg.agent_kwargs())
    obs, n = synth.make_observations(cfg, batch=64)
    act = synth.make_actions(cfg, obs, n)
    with torch.no_grad():
        logp0 = agent.step(obs, act)['logp'].cpu().numpy()
    old_logp, adv, ret = synth.make_ppo_targets(cfg, logp0)
    data = dict(obs=obs, act=act, logp=old_logp, adv=adv, ret=ret)

    def epoch(fused):
        agent.fused_ppo = fused
        agent.zero_grad()
        infos = []
        for lo, hi in ((0, 40), (40, 64)):   # two minibatches of one epoch: gradients accumulate, ONE all-reduce at the end
            loss, info = ppo.compute_loss(agent, {k: v[lo:hi] for k, v in data.items()}, 0.2, 0.5, 0.01)
            loss.backward()
            infos.append((float(loss.item()), dict(info)))
        grads = torch.cat([p.grad.reshape(-1) for p in agent.parameters()]).cpu().numpy()
        return infos, grads

    res = {'single': {f: epoch(f) for f in (True, False)}}        # not sharded yet: the single-GPU answer, on every rank
    parallel.shard_agent(agent)
    res['sharded'] = {f: epoch(f) for f in (True, False)}
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two CUDA devices')
def test_minibatch_sharded_over_two_gpus_matches_single_gpu():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for fused in (True, False):
        ref_infos, ref_grads = out[0]['single'][fused]
        gnorm = float(np.linalg.norm(ref_grads))
        a, b = out[0]['sharded'][fused], out[1]['sharded'][fused]
        assert [i for _, i in a[0]] == [i for _, i in b[0]], 'the loss info must be bit-identical on both ranks (same early-stop branch, ppo.py:138-140)'
        for (l0, i0), (l1, i1), (l2, _) in zip(ref_infos, a[0], b[0]):
            if fused:   # lazily reduced info: the loss tensor's value is the rank's share of the global loss
                assert abs(l1 + l2 - l0) <= 1e-6 * max(1.0, abs(l0))
            else:
                assert l1 == l2 and abs(l0 - l1) <= 1e-6 * max(1.0, abs(l0))
            for key in i0:
                assert abs(i0[key] - i1[key]) <= 1e-6 * max(1.0, abs(i0[key])), key
        for rank_res in (a, b):
            err = float(np.linalg.norm(rank_res[1] - ref_grads))
            assert err <= 1e-5 * gnorm, (fused, err, gnorm)
