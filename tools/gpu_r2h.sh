set -x
mkdir -p gpurun_out/r2h
cd /root/repo
for H in 256 128; do
  for TC in 0 1; do
    TAMPC_TC=$TC timeout 200 python tools/perf_sweep.py --synthetic $H --precision tf32x3 --K 1024 4096 16384 --iters 10 2>&1 | sed "s/^/H=$H TC=$TC /" >> gpurun_out/r2h/tc_smallk.jsonl
  done
done
cut -c1-250 gpurun_out/r2h/tc_smallk.jsonl
export TAMPC_TC=1
python tools/perf_sweep.py --synthetic 256 --precision tf32x3 --K 131072 --iters 3 > gpurun_out/r2h/plain256.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_mlp -s 3 -c 1 -f -o gpurun_out/r2h/tc_h256_k131072 python tools/perf_sweep.py --synthetic 256 --precision tf32x3 --K 131072 --iters 3 > gpurun_out/r2h/ncu256.log 2>&1
python tools/perf_sweep.py --synthetic 128 --precision tf32x3 --K 131072 --iters 3 > gpurun_out/r2h/plain128.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_mlp -s 3 -c 1 -f -o gpurun_out/r2h/tc_h128_k131072 python tools/perf_sweep.py --synthetic 128 --precision tf32x3 --K 131072 --iters 3 > gpurun_out/r2h/ncu128.log 2>&1
tail -3 gpurun_out/r2h/ncu256.log gpurun_out/r2h/ncu128.log
ls -la gpurun_out/r2h
