"""GPU, >= 2 devices: K-sharded command over NCCL equals the unsharded one (same injected noise), and every rank
ends with a bitwise-identical nominal sequence.  Skipped on a single-GPU box (the driver's scaling run covers N>1)."""
import os
import socket

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, precision, exchange, out):
    import torch.distributed as dist
    from tampc_b200.mppi import B200MPPI
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    z, spec = load_golden('push_nominal')
    m = B200MPPI(spec, precision=precision, device=rank, rank=rank, world_size=world, U_init=torch.from_numpy(z['in.U']),
                 exchange=exchange)
    assert m._p2p == (exchange == 'p2p')
    a1 = m.command(z['in.state'], noise=z['in.noise']).numpy()
    U1 = m.U.numpy()
    # a second command on the on-device Philox stream: the draw depends on the GLOBAL sample index only
    a2 = m.command(z['in.state']).numpy()
    for _ in range(20):  # many back-to-back commands: the double-buffered exchange never mixes sequence numbers
        m.command(z['in.state'])
    out[rank] = (a1, U1, a2, m.U.numpy(), m.stats['argmin'])
    m.close()
    dist.destroy_process_group()


@pytest.mark.parametrize('exchange', ['p2p', 'nccl'])
@pytest.mark.parametrize('precision', ['f64', 'tf32x3'])
def test_two_gpu_sharded_command(cuda_lib, precision, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from tampc_b200.mppi import B200MPPI
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, precision, exchange, out), nprocs=2, join=True)
    z, spec = load_golden('push_nominal')
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][3], out[1][3])
    tol = 1e-8 if precision == 'f64' else 1e-4
    np.testing.assert_allclose(out[0][0], z['ref.action'], rtol=0, atol=tol)
    np.testing.assert_allclose(out[0][1], z['ref.U'], rtol=0, atol=tol)
    # single-GPU run of the same two commands
    m = B200MPPI(spec, precision=precision, U_init=torch.from_numpy(z['in.U']))
    m.command(z['in.state'], noise=z['in.noise'])
    a2 = m.command(z['in.state']).numpy()
    for _ in range(20):
        m.command(z['in.state'])
    np.testing.assert_allclose(out[0][2], a2, rtol=0, atol=1e-9)
    np.testing.assert_allclose(out[0][3], m.U.numpy(), rtol=0, atol=1e-7)
    m.close()
