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
