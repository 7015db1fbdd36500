"""Host-side logic of the data-parallel path on CPU with the gloo backend (world_size 2): shard bounds, the autograd-aware
all-gather of per-shard outputs, and that sharded losses reproduce the single-process gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from molgym_b200 import parallel


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 140, 141, 1024):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans[:-1], spans[1:]):
                assert b == c and 0 <= (b - a) - (d - c) <= 1


def test_gather_shards_gloo_world2_matches_single_process():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    n, world = 11, 2
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(parallel._gloo_selftest_worker, args=(world, port, n, out), nprocs=world, join=True)
    theta = torch.tensor([0.3, -0.7], requires_grad=True)
    x = torch.linspace(-1, 1, n)
    full = torch.sin(theta[0] * x) + theta[1] * x**2
    loss = (full * torch.cos(x)).mean()
    loss.backward()
    for rank in range(world):
        l, g, f = out[rank]
        assert abs(l - loss.item()) < 1e-7
        np.testing.assert_allclose(g, theta.grad.numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(f, full.detach().numpy(), rtol=1e-6, atol=1e-7)


def test_sharded_agent_world2_matches_single_process():
    """The product's CovariantAC (on the kernel emulator) sharded over two gloo ranks: fused step and evaluate-mode step()
    return the GLOBAL loss / info on both ranks (bit-identical, so every rank takes the same early-stop branch of
    molgym/ppo.py:138-140), and after the deferred all-reduce every rank holds the single-process gradient."""
    from tests.cusim import emu_agent
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    manager = mp.Manager()
    out = manager.dict()
    mp.spawn(emu_agent.sharded_worker, args=(2, port, out), nprocs=2, join=True)
    cfg, agent, data = emu_agent.make_emu_case()
    for fused in (True, False):
        ref_infos, ref_grads = emu_agent.run_ppo_epoch(agent, data, fused)
        gnorm = float(ref_grads.norm())
        for (l0, i0), (la, ia), (lb, ib) in zip(ref_infos, out[0][fused][0], out[1][fused][0]):
            assert ia == ib                                    # the loss info is identical on both ranks, bit for bit
            if fused:   # lazily reduced info: the loss tensor's VALUE is the rank's share of the global loss (its gradient is exact)
                assert abs(la + lb - l0) <= 1e-6 * max(1.0, abs(l0))
                assert abs(ia['total_loss'] - l0) <= 1e-6 * max(1.0, abs(l0))
            else:
                assert la == lb and abs(la - l0) <= 1e-6 * max(1.0, abs(l0))
            for key in i0:
                assert abs(ia[key] - i0[key]) <= 1e-6 * max(1.0, abs(i0[key])), key
        for rank in (0, 1):
            err = float(np.linalg.norm(out[rank][fused][1] - ref_grads.numpy()))
            assert err <= 1e-5 * gnorm, (fused, rank, err, gnorm)
    for rank in (0, 1):
        pending_before, pending_after, _ = out[rank]['hook']
        assert pending_before and not pending_after
    # molgym_b200.ppo.train on the sharded agent (each rank collects only its slice of every minibatch) = the single-process run
    _, agent_t, data_t = emu_agent.make_emu_case()
    ref_info, ref_params = emu_agent.run_train(agent_t, data_t)
    for rank in (0, 1):
        info, params = out[rank]['train']
        assert info.keys() == ref_info.keys()
        for key in ref_info:
            assert abs(info[key] - ref_info[key]) <= 1e-5 * max(1.0, abs(ref_info[key])), (rank, key, info[key], ref_info[key])
        assert float(np.abs(params - ref_params).max()) <= 1e-5, rank
    assert np.array_equal(out[0]['train'][1], out[1]['train'][1])   # both ranks hold the same parameters
    np.testing.assert_array_equal(out[0]['hook'][2], out[1]['hook'][2])   # the replicas stayed in lockstep
