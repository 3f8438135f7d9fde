set -x
mkdir -p gpurun_out/r2i
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i/pytest_gpu.log
tail -6 gpurun_out/r2i/pytest_gpu.log | cut -c1-200
for H in 256 128; do
  timeout 200 python tools/perf_sweep.py --synthetic $H --precision tf32x3 --K 1024 65536 1048576 --iters 5 2>&1 | sed "s/^/H=$H /" >> gpurun_out/r2i/tc_sweep.jsonl
done
cut -c1-260 gpurun_out/r2i/tc_sweep.jsonl
