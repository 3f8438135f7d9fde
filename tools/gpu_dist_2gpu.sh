# two GPUs: sharded parity tests (kept as a log under profiles/), bench at N=2 with shard_parity + strong K=1M points
set -x
mkdir -p gpurun_out/r2e
cd /root/repo
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_gp.py -v -x > gpurun_out/r2e/pytest_dist_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e/pytest_dist_2gpu.log
tail -15 gpurun_out/r2e/pytest_dist_2gpu.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2e/bench_n2.json 2> gpurun_out/r2e/bench_n2.err; echo "bench rc=$?"
tail -5 gpurun_out/r2e/bench_n2.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2e/bench_n2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus','collective')})
print(json.dumps(l.get('shard_parity'),indent=1)[:2500])
print(json.dumps(l.get('strong_K1M'),indent=1)[:1500])
PY
