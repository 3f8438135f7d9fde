#!/bin/bash
# run-time-shaped kernel + persistent multi-wave kernel: parity, then persistent vs plain grid at several K
set -x
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 600 python -m pytest tests/test_gpu_generic.py tests/test_gpu_tc.py tests/test_gpu_synthetic.py tests/test_gpu_properties.py -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/pytest.log | cut -c1-250
for pe in 0 1; do
  TAMPC_PERSIST=$pe timeout 200 python tools/perf_sweep.py --precision tf32x3 --K 98304 131072 196608 262144 393216 524288 1048576 --iters 5 | sed "s/^/persist=$pe /" >> $O/persist_push.jsonl
  TAMPC_PERSIST=$pe timeout 200 python tools/perf_sweep.py --synthetic 32 --precision tf32x3 --K 98304 131072 196608 262144 393216 524288 1048576 --iters 5 | sed "s/^/persist=$pe /" >> $O/persist_h32.jsonl
done
cut -c1-200 $O/persist_push.jsonl $O/persist_h32.jsonl
TAMPC_GENERIC=1 timeout 200 python tools/perf_sweep.py --precision f64 mixed --K 8192 65536 --iters 5 | sed "s/^/generic /" > $O/generic_push.jsonl
timeout 200 python tools/perf_sweep.py --precision f64 mixed --K 8192 65536 --iters 5 | sed "s/^/tuned /" >> $O/generic_push.jsonl
cut -c1-200 $O/generic_push.jsonl
