#!/bin/bash
# staged weights in the run-time-shaped kernel + persistent launch as default: full GPU suite, bench, sweeps
set -x
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/pytest_gpu.log | cut -c1-250
TAMPC_GENERIC=1 timeout 200 python tools/perf_sweep.py --precision f64 mixed --K 8192 65536 --iters 5 | sed "s/^/generic /" > $O/generic_push.jsonl
timeout 200 python tools/perf_sweep.py --synthetic 48 --precision f64 --K 8192 65536 --iters 5 | sed "s/^/generic_h48 /" >> $O/generic_push.jsonl
cut -c1-220 $O/generic_push.jsonl
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-steps 1 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cut -c1-300 $O/bench.json
