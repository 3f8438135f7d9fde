# Round-2 first GPU pass (one GPU): parity tests, tcgen05 / TMEM probe, bench.
set -x
mkdir -p gpurun_out/r2a
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest_gpu.log
tail -5 gpurun_out/r2a/pytest_gpu.log
timeout 120 tools/peaks/tcgen05_probe > gpurun_out/r2a/tcgen05_probe.jsonl 2> gpurun_out/r2a/tcgen05_probe.err; echo "probe rc=$?"
cat gpurun_out/r2a/tcgen05_probe.jsonl | cut -c1-250
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2a/bench.err
cut -c1-1500 gpurun_out/r2a/bench.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2a/bench_reference.json 2>> gpurun_out/r2a/bench.err
