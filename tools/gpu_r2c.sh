set -x
mkdir -p gpurun_out/r2c
cd /root/repo
timeout 120 tools/peaks/tcgen05_issue > gpurun_out/r2c/tcgen05_issue.jsonl 2> gpurun_out/r2c/tcgen05_issue.err; echo "probe rc=$?"
cut -c26-400 gpurun_out/r2c/tcgen05_issue.jsonl
timeout 300 python -m pytest tests/test_gpu_gp.py -q -s 2>&1 | grep -E "cost rel|passed|failed" | head -30
