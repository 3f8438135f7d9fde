set -x
mkdir -p gpurun_out/r2b
cd /root/repo
timeout 120 tools/peaks/tcgen05_probe > gpurun_out/r2b/tcgen05_probe.jsonl 2> gpurun_out/r2b/tcgen05_probe.err; echo "probe rc=$?"
head -c 20000 gpurun_out/r2b/tcgen05_probe.jsonl | grep -v gemm_check | cut -c1-250
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2b/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest_gpu.log
tail -15 gpurun_out/r2b/pytest_gpu.log
