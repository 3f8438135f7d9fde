#!/bin/bash
# last check of the round: the driver's three GPU entry points (pytest -m gpu, smoke(), bench.py defaults)
set -x
mkdir -p gpurun_out/r2y
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2y/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2y/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2y/smoke.log
timeout 400 python bench.py > gpurun_out/r2y/bench_default.json 2> gpurun_out/r2y/bench_default.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/r2y/bench_default.json
