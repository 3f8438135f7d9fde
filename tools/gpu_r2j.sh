set -x
mkdir -p gpurun_out/r2j
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_synthetic.py -x -q > gpurun_out/r2j/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j/pytest_tc.log
tail -4 gpurun_out/r2j/pytest_tc.log | cut -c1-200
TAMPC_TC_ISSUERS=2 timeout 600 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/r2j/pytest_tc2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j/pytest_tc2.log
tail -4 gpurun_out/r2j/pytest_tc2.log | cut -c1-200
for NI in 1 2; do
for H in 256 128; do
  TAMPC_TC_ISSUERS=$NI timeout 200 python tools/perf_sweep.py --synthetic $H --precision tf32x3 --K 1024 65536 1048576 --iters 5 2>&1 | sed "s/^/NI=$NI H=$H /" >> gpurun_out/r2j/tc_sweep.jsonl
done
done
cut -c1-200 gpurun_out/r2j/tc_sweep.jsonl
