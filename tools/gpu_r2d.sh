set -x
mkdir -p gpurun_out/r2d
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_integration.py tests/test_gpu_gp.py -q -s -x 2>&1 | grep -E "cost rel|top-2|passed|failed|Error|error|assert" | head -60
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2d/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest_gpu.log
tail -8 gpurun_out/r2d/pytest_gpu.log
