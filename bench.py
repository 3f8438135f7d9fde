#!/usr/bin/env python
"""bench.py — MPPI sample-steps/s and command() latency of the B200 path, beside the CPU reference path.

Contract (driver):  python bench.py --gpus N --steps K --warmup W [--impl reference]
  N>1 is launched under torchrun, one rank per GPU; rank 0 prints ONE JSON line.

Workload (BASELINE.json configs[1]): block-pushing TAMPC — RexExtract latent dynamics from
the committed checkpoints (packaged as tests/golden/push_nominal.npz), QR goal cost + 3-trap set cost + terminal
x50, T=40, lambda=0.01, K=8192 samples per GPU; synthetic start state; on-device Philox noise.
A "step" is one MPPI command(): K*T sample-steps.
  value  : sample-steps/s, device-timed (CUDA events on the launching stream), state already resident in HBM
  e2e    : the same through the public API B200MPPI.command(host state) -> host action (H2D + D2H + sync inside)
  N>1    : weak scaling, K=8192 per GPU, samples sharded, one all_gather of the softmax partials per command
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_PER_GPU = 8192
METRIC = "mppi_sample_steps_per_sec"
UNIT = "sample-steps/s"
FP32_PEAK_TFLOPS = 74.1  # tools/peaks/fma_peaks ffma2_reg on this pool's B200 (profiles/fma_peaks_r1.jsonl)
FP64_TENSOR_PEAK_TFLOPS = 37.1  # tools/peaks/mma_peaks dmma (profiles/mma_peaks_r1.jsonl)
TF32_MMA_PEAK_TFLOPS = 278.0  # tools/peaks/mma_peaks mma.sync m16n8k8 tf32
F16_MMA_PEAK_TFLOPS = 556.0  # tools/peaks/mma_peaks mma.sync m16n8k16 f16
# DRAM traffic of one rollout launch at K=8192 from the committed ncu capture (the kernel reads its ~45 KB image and
# the nominal sequence, writes 8*K bytes of costs that stay in L2): compute-bound by four orders of magnitude
NCU_DRAM_BYTES_PER_LAUNCH = {'tf32x3': 75776, 'f16x3': 70912}


# The driver's line is the default workload (BASELINE.json configs[1]).  The others exist to put the remaining configs
# on record with the same harness (profiles/): configs[2] peg, configs[0] pendulum, configs[3]/[4] the synthetic sweep's
# 1M-sample points, sharded over the GPUs (strong scaling: K is the total).
WORKLOADS = {
    'push': dict(golden='push_nominal', K_per_gpu=K_PER_GPU, name='push_rex_extract_K8192_T40_P3'),
    'peg': dict(golden='peg_nominal', K_per_gpu=32768, name='peg_rex_extract_K32768_T10_P4'),
    'pendulum': dict(golden='pendulum', K_per_gpu=100, name='pendulum_analytic_K100_T15'),
    'synth_h32': dict(synthetic=32, K_total=1 << 20, name='synthetic_mlp_H32_K1M_T40_P3'),
    'synth_h256': dict(synthetic=256, K_total=1 << 20, name='synthetic_mlp_H256_K1M_T40_P3'),
}


def load_spec(workload, world, K=None):
    """(arrays with in.state / in.U, RolloutSpec) of a workload at `world` GPUs (K overrides the sample count)."""
    from tampc_b200 import spec as S
    w = WORKLOADS[workload]
    if 'synthetic' in w:
        from tampc_b200.synthetic import make_synthetic_spec
        spec, x0 = make_synthetic_spec(H=w['synthetic'], K=K or w['K_total'], T=40)
        return {'in.state': x0, 'in.U': np.zeros((40, 3))}, spec
    z = np.load(os.path.join(ROOT, 'tests', 'golden', w['golden'] + '.npz'))
    spec = S.spec_from_arrays(z)
    spec.sampler.K = K or w['K_per_gpu'] * world
    return z, spec


def load_push_spec(K):
    return load_spec('push', 1, K)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([p.strip() for p in line.split(',')])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_step(spec, z, state, U, seed):
    """One command() of the CPU oracle port on a fresh noise draw; returns seconds."""
    import torch
    from oracle import mppi_oracle as O
    noise = O.sample_noise(spec, seed)
    t0 = time.perf_counter()
    O.command(spec, torch.from_numpy(state), U, noise)
    return time.perf_counter() - t0


def torch_cuda_f64_comparator(spec, z, state, U, steps=3):
    """Secondary comparator: the reference's eager float64 torch op sequence with its tensors on the
    GPU — what the authors actually ran (tampc/util.py:8 get_device()).  Same oracle port, device 'cuda'."""
    import torch
    from oracle import mppi_oracle as O
    O.DEVICE = 'cuda'
    try:
        times = []
        for i in range(steps + 1):
            noise = O.sample_noise(spec, 300 + i)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = O.command(spec, torch.from_numpy(state), U, noise)
            float(out['action'][0])  # the D2H read of the action, like MPC.command's u.cpu().numpy()
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        return float(np.mean(times[1:]))
    finally:
        O.DEVICE = 'cpu'


def cpu_sample_K(workload):
    """Sample count of one CPU command: the workload's own K per GPU, bounded so that a command stays ~0.1-0.5 s."""
    w = WORKLOADS[workload]
    return min(w.get('K_per_gpu', 8192), 32768)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU torch path (oracle port: /root/reference does not exist on the GPU box),
    float64, all host threads, same config/metric; each step is one full K=8192 command."""
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    Kc = cpu_sample_K(args.workload)
    z, spec = load_spec(args.workload, 1, Kc)
    state, U = z['in.state'].copy(), torch.from_numpy(z['in.U'].copy())
    for i in range(args.warmup):
        cpu_reference_step(spec, z, state, U, 100 + i)
    times = [cpu_reference_step(spec, z, state, U, 200 + i) for i in range(args.steps)]
    T = spec.sampler.T
    ms = 1e3 * float(np.mean(times))
    val = Kc * T / (ms * 1e-3)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]['name'], "K": Kc, "T": T, "rollout_replicas_M": 1,
                   "note": "deterministic dynamics: one replica; the reference scripts run M=10 identical replicas"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "{} full commands of K={} T={} (oracle/mppi_oracle.py, torch float64)".format(args.steps, Kc, T)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "latency_ms_p50": 1e3 * float(np.median(times)), "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('TAMPC_BENCH_PRECISION', 'tf32x3'))
    ap.add_argument('--cpu-steps', type=int, default=12, help='commands of the CPU baseline sample')
    ap.add_argument('--workload', default='push', choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from tampc_b200.mppi import B200MPPI

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    z, spec = load_spec(args.workload, world)
    K = spec.sampler.K
    strong = 'K_total' in WORKLOADS[args.workload]
    T = spec.sampler.T
    state = z['in.state'].copy()
    mppi = B200MPPI(spec, precision=args.precision, device=local_rank, rank=rank, world_size=world, seed=1234,
                    U_init=torch.from_numpy(np.asarray(z['in.U']).copy()), use_torch_stream=True, max_horizon=max(T, 64))
    stream = mppi.torch_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-timed throughput (state resident, L2 flushed between steps)
    if world > 1:
        mppi.set_state_resident(state)

    def device_step():
        if world == 1 or mppi._p2p:
            mppi.command_resident()  # sharded: the merge kernel exchanges the partials over peer memory itself
        else:
            mppi.command(state)      # NCCL fallback: the all-gather is issued from the host

    if world == 1:
        mppi.set_state_resident(state)
    for _ in range(args.warmup):
        device_step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = mppi.stats['launches']
    wall0 = time.perf_counter()
    for e0, e1 in ev:
        with torch.cuda.stream(stream):
            flush.zero_()
            e0.record()
            device_step()
            e1.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = mppi.stats['launches'] - launches0  # kernels of this library launched inside the timed region
    dev_ms = np.array([e0.elapsed_time(e1) for e0, e1 in ev])
    ms_per_step = float(dev_ms.mean())

    # ---------------- rollout kernel alone (roofline of the dominant kernel)
    roll_ms = None
    if world == 1:
        for _ in range(3):
            mppi.rollout_resident()
        torch.cuda.synchronize()
        evr = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 50))]
        for e0, e1 in evr:
            with torch.cuda.stream(stream):
                flush.zero_()
                e0.record()
                mppi.rollout_resident()
                e1.record()
        torch.cuda.synchronize()
        roll_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evr]))

    # ---------------- end to end through the public API: host state in, host action out
    for _ in range(args.warmup):
        mppi.command(state)
    barrier()
    lat = []
    t_all0 = time.perf_counter()
    for i in range(args.steps):
        t0 = time.perf_counter()
        a = mppi.command(state)
        lat.append(time.perf_counter() - t0)
    barrier()
    e2e_wall = time.perf_counter() - t_all0
    # The timed regions last milliseconds, shorter than nvidia-smi's sampling period: keep the identical load running
    # for another half second so that the clock / throttle record has samples taken under this load.
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 0.6:
        for _ in range(50):
            device_step()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    if clk is not None:
        clk["note"] = "sampled every 100 ms from the first timed step until 0.6 s of the same load after the last one"

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_per_step, e2e_wall, float(np.median(lat))], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_per_step, e2e_wall, lat_p50 = (float(v) for v in t.tolist())
    else:
        lat_p50 = float(np.median(lat))

    if rank == 0:
        value = K * T / (ms_per_step * 1e-3)
        e2e_val = K * T * args.steps / e2e_wall
        flop_per_step = 2.0 * spec.dynamics.macs_per_step  # 7046 for push
        roofline = None
        if roll_ms is not None:
            achieved = K * T * flop_per_step / (roll_ms * 1e-3) * 1e-12
            # the kernel is compute-bound (it moves ~100 KB of DRAM traffic per launch); which pipe bounds it
            # depends on the arithmetic: every peak below was measured on this pool's B200 with tools/peaks
            # (MEASURED_PEAKS.json holds only the HBM copy and the cuBLAS bf16 tcgen05 GEMM)
            if spec.dynamics.kind == 2:  # analytic pendulum: no network layers, scalar FP32/FP64 arithmetic only
                bound, peak = "fp32_fma", FP32_PEAK_TFLOPS
                src = "FFMA2 register-operand peak measured (profiles/fma_peaks_r1.jsonl); ~20 FLOP + 1 sin per sample-step"
            elif args.precision == 'tf32x3':
                bound, peak = "tensor", TF32_MMA_PEAK_TFLOPS / 3.0
                src = ("mma.sync m16n8k8 tf32 measured {:.0f} TFLOP/s (profiles/mma_peaks_r1.jsonl) / 3 products per "
                       "algorithmic MAC (error-compensated 3xTF32 split)").format(TF32_MMA_PEAK_TFLOPS)
            elif args.precision == 'f16x3':
                bound, peak = "tensor", F16_MMA_PEAK_TFLOPS / 3.0
                src = ("mma.sync m16n8k16 f16 measured {:.0f} TFLOP/s (profiles/mma_peaks_r1.jsonl) / 3 products per "
                       "algorithmic MAC (error-compensated 3xFP16 split)").format(F16_MMA_PEAK_TFLOPS)
            elif args.precision == 'f64':
                bound, peak = "tensor", FP64_TENSOR_PEAK_TFLOPS
                src = "mma.sync m8n8k4 f64 (DMMA) measured (profiles/mma_peaks_r1.jsonl)"
            else:
                bound, peak = "fp32_fma", FP32_PEAK_TFLOPS
                src = "FFMA2 register-operand peak measured (profiles/fma_peaks_r1.jsonl)"
            roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak,
                        "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(args.precision) if args.workload == 'push' else None,
                        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size, one "
                                          "ncu --set full capture (profiles/r1h_rollout_split3_tf32x3_k8192_ncu_summary.txt; f16x3: r1e_rollout_split_f16x3_k8192_ncu_summary.txt)",
                        "kernel": "rollout (fused K x T): rollout_mma_split_kernel at this K (three warps per 16-sample "
                                  "tile: chain | decoder | cost), rollout_mma_kernel above K=12288",
                        "kernel_ms": roll_ms, "algorithmic_flop_per_launch": K * T * flop_per_step,
                        "peak_source": src}
        # CPU baseline: bounded sample of the same workload on the host cores (oracle port, float64)
        torch.set_num_threads(os.cpu_count() or 1)
        Kc = cpu_sample_K(args.workload)
        zc, spec_c = load_spec(args.workload, 1, Kc)
        Uc = torch.from_numpy(np.asarray(z['in.U']).copy())
        cpu_reference_step(spec_c, zc, state, Uc, 1)
        cpu_t = [cpu_reference_step(spec_c, zc, state, Uc, 2 + i) for i in range(args.cpu_steps)]
        cpu_val = Kc * T / float(np.mean(cpu_t))
        cuda_t = torch_cuda_f64_comparator(spec_c, zc, state, Uc) if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": {"f64": "f64", "mixed": "f32 networks, f64 state+cost", "f32": "f32",
                      "tf32x3": "3xTF32 (f32-equivalent) networks, f32 state, f64 cost sum",
                      "f16x3": "3xFP16 (f32-equivalent) networks, f32 state, f64 cost sum"}.get(args.precision, args.precision),
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]['name'], "K_per_gpu": mppi.K_local, "K": K, "T": T,
                       "precision": args.precision, "noise": "on-device Philox4x32-10", "rollout_replicas_M": 1,
                       "l2": "flushed between timed steps (256 MiB memset on the same stream, outside the events)",
                       "parallelism": "samples sharded over {} GPU(s)".format(world)},
            "latency_ms_p50": lat_p50 * 1e3,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(spec.dynamics.nx * 8),
                    "d2h_bytes_per_step": int(spec.sampler.nu * 8 + 24)},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "{} commands of K={} T={} (oracle/mppi_oracle.py, torch float64, M=1)".format(
                                 args.cpu_steps, Kc, T),
                             "torch_cuda_f64": None if cuda_t is None else {
                                 "value": Kc * T / cuda_t, "unit": UNIT, "ms_per_command": 1e3 * cuda_t,
                                 "note": "secondary comparator: the same eager float64 torch op sequence with its tensors "
                                         "on this GPU (what the reference's authors ran), 3 commands"}},
            "wall_s_timed_region": wall,
        }
        print(json.dumps(line), flush=True)
    mppi.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
