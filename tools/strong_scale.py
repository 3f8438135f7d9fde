"""K = 1M sharded over the ranks of one node (BASELINE.json configs[3]/[4]): bench.py's `strong_K1M` points alone, without
the CPU legs.  torchrun --nproc-per-node N tools/strong_scale.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import bench  # noqa: E402
from tampc_b200.mppi import B200MPPI  # noqa: E402


def main():
    rank, local_rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = bench.strong_scaling_points(None, rank, local_rank, world, dist, torch, B200MPPI, flush, barrier)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "strong_K1M": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
