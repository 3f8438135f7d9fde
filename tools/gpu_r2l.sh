set -x
mkdir -p gpurun_out/r2l
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_synthetic.py -q > gpurun_out/r2l/pytest_tc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l/pytest_tc.log
tail -5 gpurun_out/r2l/pytest_tc.log | cut -c1-200
timeout 300 python tools/tc_timeline.py --H 256 --K 1024 > gpurun_out/r2l/timeline.txt 2>&1
timeout 300 python tools/tc_timeline.py --H 128 --K 1024 >> gpurun_out/r2l/timeline.txt 2>&1
for H in 256 128; do
  timeout 200 python tools/perf_sweep.py --synthetic $H --precision tf32x3 --K 1024 65536 1048576 --iters 5 2>&1 | sed "s/^/H=$H /" >> gpurun_out/r2l/tc_sweep.jsonl
done
