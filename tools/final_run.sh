# End-of-round measurement pass (one GPU): bench arms, workloads, K sweep, ncu launch list and --set full captures.
#   gpurun -- bash tools/final_run.sh ; outputs under gpurun_out/r1h/, summaries copied to profiles/r1h_*
set -x
mkdir -p gpurun_out/r1h
cd /root/repo
timeout 300 python bench.py > gpurun_out/r1h/bench.json 2> gpurun_out/r1h/bench.err
timeout 300 python bench.py --impl reference > gpurun_out/r1h/bench_reference.json 2>> gpurun_out/r1h/bench.err
timeout 300 python bench.py --precision f16x3 > gpurun_out/r1h/bench_f16x3.json 2>> gpurun_out/r1h/bench.err
for w in peg pendulum synth_h32 synth_h256; do timeout 300 python bench.py --workload $w --cpu-steps 2 2>> gpurun_out/r1h/bench.err | tail -1 >> gpurun_out/r1h/workloads_n1.jsonl; done
timeout 200 python tools/perf_sweep.py --precision tf32x3 --K 500 4736 8192 12288 65536 75776 262144 1048576 > gpurun_out/r1h/sweep_tf32x3.jsonl 2>&1
timeout 200 python tools/perf_sweep.py --precision f16x3 --K 8192 65536 1048576 > gpurun_out/r1h/sweep_f16x3.jsonl 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1h/launches.csv python bench.py --steps 20 --warmup 3 --cpu-steps 1 > gpurun_out/r1h/ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_mma_split -s 5 -c 1 -f -o gpurun_out/r1h/split3_k8192 python tools/perf_sweep.py --precision tf32x3 --K 8192 --iters 10 > gpurun_out/r1h/ncu_split.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_mma_mix -s 5 -c 1 -f -o gpurun_out/r1h/mix_k65536 python tools/perf_sweep.py --precision tf32x3 --K 65536 --iters 10 > gpurun_out/r1h/ncu_mix.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:partials_combine -s 5 -c 1 -f -o gpurun_out/r1h/tail_k8192 python tools/perf_sweep.py --precision tf32x3 --K 8192 --iters 10 > gpurun_out/r1h/ncu_tail.log 2>&1
ls -la gpurun_out/r1h
tail -2 gpurun_out/r1h/bench.json | cut -c1-400
