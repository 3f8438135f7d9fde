# Round-2 measurement pass (one GPU): tests, bench arms, workloads, K sweeps, ncu launch list and --set full captures.
#   gpurun -- bash tools/final_run_r2.sh ; outputs under gpurun_out/r2z/, summaries copied to profiles/r2_*
set -x
rm -rf gpurun_out/r2z
mkdir -p gpurun_out/r2z
cd /root/repo
O=gpurun_out/r2z
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_reference.json 2>> $O/bench.err
for w in peg pendulum synth_h32 synth_h256; do timeout 300 python bench.py --workload $w --cpu-steps 2 --no-extras 2>> $O/bench.err | tail -1 >> $O/workloads_n1.jsonl; done
timeout 300 python tools/perf_sweep.py --precision tf32x3 --K 500 4736 8192 12288 65536 131072 262144 1048576 > $O/sweep_push_tf32x3.jsonl 2>&1
for H in 32 64 128 256; do timeout 200 python tools/perf_sweep.py --synthetic $H --precision tf32x3 --K 1024 16384 65536 131072 262144 1048576 --iters 5 >> $O/sweep_synth.jsonl 2>&1; done
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
python bench.py --steps 20 --warmup 3 --cpu-steps 1 --no-extras > $O/plain_bench.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/bench_launches.csv python bench.py --steps 20 --warmup 3 --cpu-steps 1 --no-extras > $O/ncu_bench.log 2>&1
# the dominant kernel of the bench (role-split rollout, K=8192) and the tcgen05 kernel at H=256
python tools/perf_sweep.py --precision tf32x3 --K 8192 --iters 10 > $O/plain_split.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_mma_split -s 5 -c 1 -f -o $O/split3_k8192 python tools/perf_sweep.py --precision tf32x3 --K 8192 --iters 10 > $O/ncu_split.log 2>&1
python tools/perf_sweep.py --synthetic 256 --precision tf32x3 --K 131072 --iters 3 > $O/plain_tc256.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_mlp -s 3 -c 1 -f -o $O/tc_h256_k131072 python tools/perf_sweep.py --synthetic 256 --precision tf32x3 --K 131072 --iters 3 > $O/ncu_tc256.log 2>&1
python tools/perf_sweep.py --precision tf32x3 --K 131072 --iters 5 > $O/plain_persist.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rollout_mma_persist -s 3 -c 1 -f -o $O/persist_k131072 python tools/perf_sweep.py --precision tf32x3 --K 131072 --iters 5 > $O/ncu_persist.log 2>&1
TAMPC_GENERIC=1 timeout 200 python tools/perf_sweep.py --precision f64 mixed --K 8192 65536 --iters 5 > $O/sweep_generic.jsonl 2>&1
timeout 200 python tools/tc_timeline.py --H 256 --K 1024 1048576 > $O/tc_timeline.txt 2>&1
timeout 200 python tools/tc_timeline.py --H 128 --K 1024 1048576 >> $O/tc_timeline.txt 2>&1
ls -la $O
cut -c1-300 $O/bench.json
