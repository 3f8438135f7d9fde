"""CPU, world_size=2 over gloo: the K-sharding host logic — shard bounds, one all_gather of per-rank partials laid
out as tampc_partial, rank-order combine — reproduces the unsharded command.  (The CUDA partial/combine kernels
implement the same algebra; on one GPU they are checked against the goldens, on N GPUs by tests/test_gpu_dist.py.)"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden
from oracle import mppi_oracle as O
from tampc_b200 import _lib as L
from tampc_b200.mppi import shard_bounds


def test_shard_bounds_cover_every_sample_once():
    for K in (1, 7, 100, 8192, 1048576 + 3):
        for world in (1, 2, 3, 4, 8):
            if K < world:
                continue
            spans = [shard_bounds(K, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == K
            for (lo, n), (lo2, _) in zip(spans, spans[1:]):
                assert lo + n == lo2 and n >= 1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    z, spec = load_golden('push_nominal')
    T, nu = spec.sampler.T, spec.sampler.nu
    U = O.shift_nominal(spec, torch.from_numpy(z['in.U']).clone())
    lo, n = shard_bounds(spec.sampler.K, rank, world)
    noise = torch.from_numpy(z['in.noise'])[lo:lo + n]
    perturbed, npost, acost = O.perturb(spec, U, noise)
    cost = O.rollout_costs(spec, torch.from_numpy(z['in.state']), perturbed) + torch.sum(perturbed * acost, dim=(1, 2))
    p = O.shard_partial(spec, cost, npost, lo)
    mine = torch.zeros(L.PARTIAL_DOUBLES, dtype=torch.double)  # tampc_partial layout: beta, eta, argmin, count, wsum
    mine[0], mine[1] = p['beta'], p['eta']
    mine[2:4] = torch.tensor([p['argmin'], n], dtype=torch.int64).view(torch.double)
    mine[4:4 + T * nu] = p['wsum'].reshape(-1)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    parts = [{'beta': float(g[0]), 'eta': float(g[1]), 'argmin': int(g[2:3].view(torch.int64)[0]),
              'wsum': g[4:4 + T * nu].view(T, nu)} for g in gathered]
    res = O.combine_partials(spec, parts, U)
    out[rank] = (res['action'].numpy(), res['U'].numpy(), res['argmin'])
    dist.destroy_process_group()


def test_two_rank_sharded_command_equals_unsharded():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    z, spec = load_golden('push_nominal')
    np.testing.assert_array_equal(out[0][1], out[1][1])  # identical on both ranks, bit for bit
    assert out[0][2] == out[1][2] == int(z['ref.argmin'])
    np.testing.assert_allclose(out[0][0], z['ref.action'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(out[0][1], z['ref.U'], rtol=0, atol=1e-10)
