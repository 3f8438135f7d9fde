set -x
mkdir -p gpurun_out/r2g
cd /root/repo
for H in 256 128 64 32; do
  for TC in 0 1; do
    TAMPC_TC=$TC timeout 200 python tools/perf_sweep.py --synthetic $H --precision tf32x3 --K 65536 1048576 --iters 5 2>&1 | sed "s/^/H=$H TC=$TC /" >> gpurun_out/r2g/tc_sweep.jsonl
  done
done
cut -c1-420 gpurun_out/r2g/tc_sweep.jsonl
