set -x
mkdir -p gpurun_out/r2f
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -s > gpurun_out/r2f/pytest_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2f/pytest_tc.log
grep -E "cost rel|passed|failed|Error|rc=|assert|^E " gpurun_out/r2f/pytest_tc.log | head -40
