#!/bin/bash
# 8-GPU box: driver-style scaling runs of bench.py (N = 8, 4, 2), sharded parity tests
set -x
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
nvidia-smi -L | head -8
timeout 300 python -m pytest tests/test_gpu_dist.py -q > $O/pytest_dist.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_dist.log
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510+n)) bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_n$n.json 2> $O/bench_n$n.err
  echo "bench n=$n rc=$?"; tail -1 $O/bench_n$n.json | cut -c1-300
done
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n=1 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err; echo "ref n=8 rc=$?"
tail -2 $O/bench_n8.err
