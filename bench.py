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
  N>1    : weak scaling, K=8192 per GPU, samples sharded; the only exchange per command is one ~1 KB softmax partial per
           rank, stored straight into every peer's CUDA-IPC mapped buffer by the merge kernel (NCCL is the bootstrap and
           the timing all-reduce only)
Extra keys on the same line (none of them replaces a contract key):
  shard_parity (N>1)  one Philox command sharded over the N ranks against the same command unsharded on rank 0, checked
                      OUTSIDE the timed region: bitwise-identical U on every rank, |U - U_unsharded|, argmin, per-sample costs;
                      the run FAILS (rc 1) if it does not hold
  strong_K1M          BASELINE.json configs[4]: K = 1,048,576 samples TOTAL, T=40, synthetic MLP H=32 and H=256, sharded over
                      the N ranks (strong scaling; N=1 is the single-GPU point of configs[3])
  f64 (N=1)           the same workload in the reference's own arithmetic (float64, DMMA kernels)
  reference_M10 (N=1) the CPU path doing the reference scripts' literal work: M=10 identical replicas of every sample
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
TCGEN05_TF32_PEAK_TFLOPS = 982.7  # tools/peaks/tcgen05_probe (profiles/tcgen05_probe_r2.jsonl): tcgen05.mma kind::tf32, M=128, N=256, A in TMEM
# DRAM traffic of one rollout launch at K=8192 from the committed ncu capture (the kernel reads its ~45 KB image and
# the nominal sequence, writes 8*K bytes of costs that stay in L2): compute-bound by four orders of magnitude
NCU_DRAM_BYTES_PER_LAUNCH = {'tf32x3': 91648, 'f16x3': 70912}


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


def make_config(workload, world, K, T):
    """The workload as both arms print it (identical dicts: the driver compares them)."""
    return {"workload": WORKLOADS[workload]['name'], "K": int(K), "K_per_gpu": int(-(-K // world)), "T": int(T),
            "rollout_replicas_M": 1, "noise": "Philox4x32-10 on the device / torch.randn on the host, seed 1234",
            "l2": "flushed between timed steps (256 MiB memset on the same stream, outside the events)",
            "parallelism": "samples sharded over {} GPU(s)".format(world)}


def cpu_reference_step(spec, z, state, U, seed):
    """One command() of the CPU oracle port on a fresh noise draw; returns seconds."""
    import torch
    from oracle import mppi_oracle as O
    noise = O.sample_noise(spec, seed)
    t0 = time.perf_counter()
    O.command(spec, torch.from_numpy(state), U, noise)
    return time.perf_counter() - t0


def cpu_reference_m10(workload, state, U, steps=2):
    """The reference scripts' literal work (push_main.py:110, peg_main.py:103: rollout_samples=10): every sample rolled out
    M=10 times although the dynamics are deterministic.  Sample-steps/s still counts unique (k, t)."""
    import torch
    from oracle import mppi_oracle as O
    Kc = cpu_sample_K(workload)
    z, spec = load_spec(workload, 1, Kc)
    spec.sampler.rollout_samples = 20 if workload == 'pendulum' else 10
    O.LITERAL_REPLICAS = True
    try:
        cpu_reference_step(spec, z, state, U, 50)
        t = [cpu_reference_step(spec, z, state, U, 51 + i) for i in range(steps)]
    finally:
        O.LITERAL_REPLICAS = False
    return {"value": Kc * spec.sampler.T / float(np.mean(t)), "unit": UNIT, "ms_per_command": 1e3 * float(np.mean(t)),
            "rollout_replicas_M": spec.sampler.rollout_samples, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "{} commands of K={} T={} at M={} (oracle/mppi_oracle.py LITERAL_REPLICAS, torch float64)".format(
                steps, Kc, spec.sampler.T, spec.sampler.rollout_samples)}


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


def cpu_sample_K(workload, world=1):
    """Sample count of one CPU command: the workload's K at `world` GPUs (weak scaling: K_per_gpu * world), bounded so
    that a command stays below ~2 s on the host cores."""
    w = WORKLOADS[workload]
    return min(w.get('K_per_gpu', 8192) * world, 131072)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU torch path (oracle port: /root/reference does not exist on the GPU box),
    float64, all host threads, the product arm's config / metric; each step is one full command of K samples."""
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    strong = 'K_total' in WORKLOADS[args.workload]
    K_cfg = WORKLOADS[args.workload]['K_total'] if strong else WORKLOADS[args.workload]['K_per_gpu'] * world
    Kc = min(K_cfg, 131072)  # a bounded sample of the workload (sample-steps/s is K-normalised)
    z, spec = load_spec(args.workload, 1, Kc)
    state, U = z['in.state'].copy(), torch.from_numpy(np.asarray(z['in.U']).copy())
    for i in range(args.warmup):
        cpu_reference_step(spec, z, state, U, 100 + i)
    times = [cpu_reference_step(spec, z, state, U, 200 + i) for i in range(args.steps)]
    T = spec.sampler.T
    ms = 1e3 * float(np.mean(times))
    val = Kc * T / (ms * 1e-3)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(args.workload, world, K_cfg, T),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "{} full commands of K={} T={} at M=1 (oracle/mppi_oracle.py, torch float64; deterministic "
                                   "dynamics: the reference's M identical replicas are rolled out once)".format(args.steps, Kc, T)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "latency_ms_p50": 1e3 * float(np.median(times)), "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def _numpy_state(z):
    return np.ascontiguousarray(np.asarray(z['in.state'], dtype=np.float64)).copy()


def shard_parity_check(precision, spec_world, z, rank, local_rank, world, dist, torch, B200MPPI, u_tol, cost_rtol):
    """One Philox command sharded over the ranks against the same command unsharded (rank 0), outside any timed region.
    The Philox counter is keyed by the GLOBAL sample index, so both draw the same K x T perturbations
    (tampc/controller/controller.py:383-402 is the semantics the merge must preserve).  In 'f64' every kernel family
    computes a sample's cost to ~1e-13, so U must agree to 1e-9; in the float32 modes a shard of K/N samples may run a
    different kernel family than the unsharded K (different float32 summation order inside a sample's trap-set cost), so
    costs agree to float32 rounding and U to the north-star action tolerance."""
    import copy
    from tampc_b200.mppi import shard_bounds
    state = _numpy_state(z)
    U0 = torch.from_numpy(np.asarray(z['in.U']).copy())
    T = spec_world.sampler.T
    K = spec_world.sampler.K
    m = B200MPPI(copy.deepcopy(spec_world), precision=precision, device=local_rank, rank=rank, world_size=world, seed=4321,
                 U_init=U0.clone(), use_torch_stream=True, max_horizon=max(T, 64))
    a = m.command(state)
    U = m.U.to('cuda')
    st = m.stats
    cost = m.cost_total.to('cuda')
    exchange = 'p2p' if m._p2p else 'nccl'
    kname = m.kernel_name
    m.close()
    allU = [torch.empty_like(U) for _ in range(world)]
    dist.all_gather(allU, U)
    # per-sample costs of every shard -> rank 0 (shards differ in size by at most one sample: pad to the largest)
    nmax = max(shard_bounds(K, r, world)[1] for r in range(world))
    pad = torch.full((nmax,), float('nan'), dtype=torch.double, device='cuda')
    pad[:cost.numel()] = cost
    allc = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(allc, pad)
    arg = torch.tensor([st['argmin']], dtype=torch.int64, device='cuda')
    allarg = [torch.empty_like(arg) for _ in range(world)]
    dist.all_gather(allarg, arg)
    out = None
    if rank == 0:
        ref = B200MPPI(copy.deepcopy(spec_world), precision=precision, device=local_rank, seed=4321, U_init=U0.clone(),
                       max_horizon=max(T, 64))
        a1 = ref.command(state)
        U1 = ref.U
        c1 = ref.cost_total
        arg1 = ref.stats['argmin']
        kname1 = ref.kernel_name
        ref.close()
        bitwise = all(bool(torch.equal(allU[0], u)) for u in allU[1:])
        csh = torch.cat([allc[r][:shard_bounds(K, r, world)[1]] for r in range(world)]).cpu()
        args_sh = [int(x) for x in allarg]
        s2 = torch.sort(c1).values
        gap = float((s2[1] - s2[0]) / s2[0].abs()) if K > 1 else float('inf')
        out = {"precision": precision, "U_bitwise_equal_across_ranks": bitwise,
               "max_abs_vs_unsharded": float((allU[0].cpu() - U1).abs().max()), "U_tolerance": u_tol,
               "action_abs_vs_unsharded": float((a - a1).abs().max()),
               "argmin_equal": bool(all(x == int(arg1) for x in args_sh)), "argmin": int(arg1), "argmin_sharded": args_sh[0],
               "top2_cost_gap_rel": gap,
               "cost_max_rel_vs_unsharded": float(((csh - c1).abs() / c1.abs()).max()), "cost_rtol": cost_rtol,
               "K": int(K), "ranks": world, "exchange": exchange, "noise": "Philox (global sample index)",
               "kernel_sharded": kname, "kernel_unsharded": kname1}
        # the argmin may only differ where the two lowest costs are closer than the arithmetic tolerance (a tie)
        argmin_ok = out["argmin_equal"] or gap < cost_rtol
        out["ok"] = bool(bitwise and argmin_ok and out["max_abs_vs_unsharded"] < u_tol and out["cost_max_rel_vs_unsharded"] < cost_rtol)
    return out


def timed_commands(mppi, state, steps, warmup, flush, barrier, torch):
    """ms per command, device-timed with CUDA events on the handle's stream, L2 flushed between commands."""
    stream = mppi.torch_stream
    mppi.set_state_resident(state)
    step = mppi.command_resident if (mppi.world_size == 1 or mppi._p2p) else (lambda: mppi.command(state))
    for _ in range(warmup):
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in ev:
        with torch.cuda.stream(stream):
            flush.zero_()
            e0.record()
            step()
            e1.record()
    barrier()
    return np.array([e0.elapsed_time(e1) for e0, e1 in ev])


def strong_scaling_points(args, rank, local_rank, world, dist, torch, B200MPPI, flush, barrier):
    """BASELINE.json configs[3]/[4]: K = 1M samples in total, T = 40, synthetic MLP of hidden width 32 and 256, sharded over
    the ranks.  Returns {workload: {...}} on rank 0 (max over ranks of the per-rank mean)."""
    out = {}
    for wl, steps in (('synth_h32', 10), ('synth_h256', 4)):
        z, spec = load_spec(wl, world)
        K, T = spec.sampler.K, spec.sampler.T
        m = B200MPPI(spec, precision='tf32x3', device=local_rank, rank=rank, world_size=world, seed=1234,
                     U_init=torch.from_numpy(np.asarray(z['in.U']).copy()), use_torch_stream=True, max_horizon=64)
        ms = float(timed_commands(m, _numpy_state(z), steps, 3, flush, barrier, torch).mean())
        kname = m.kernel_name
        m.close()
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        flop = 2.0 * spec.dynamics.macs_per_step
        tf = K * T * flop / (ms * 1e-3) * 1e-12
        peak = (TCGEN05_TF32_PEAK_TFLOPS if 'tcgen05' in kname else TF32_MMA_PEAK_TFLOPS) / 3.0 * world
        out[wl] = {"workload": WORKLOADS[wl]['name'], "K_total": int(K), "K_per_gpu": int(-(-K // world)), "T": int(T), "n_gpus": world,
                   "scaling": "strong", "steps": steps, "warmup": 3, "ms_per_step": ms, "value": K * T / (ms * 1e-3), "unit": UNIT,
                   "tflops_algorithmic": tf, "kernel": kname,
                   "frac_of_tensor_peak": tf / peak,
                   "peak_note": "whole command (rollout + softmax update) against {} GPU(s) x the measured {} peak / 3 products per MAC".format(
                       world, "tcgen05 kind::tf32 982.7 TFLOP/s" if 'tcgen05' in kname else "mma.sync tf32 278 TFLOP/s")}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default=os.environ.get('TAMPC_BENCH_PRECISION', 'tf32x3'))
    ap.add_argument('--cpu-steps', type=int, default=12, help='commands of the CPU baseline sample')
    ap.add_argument('--workload', default='push', choices=sorted(WORKLOADS))
    ap.add_argument('--no-extras', action='store_true', help='skip strong_K1M / f64 / reference_M10 (profiling runs)')
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
    state = _numpy_state(z)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- sharded == unsharded, before anything is timed (N > 1)
    parity = None
    if world > 1:
        p64 = shard_parity_check('f64', spec, z, rank, local_rank, world, dist, torch, B200MPPI, u_tol=1e-9, cost_rtol=1e-9)
        barrier()
        pb = None
        if args.precision != 'f64':
            pb = shard_parity_check(args.precision, spec, z, rank, local_rank, world, dist, torch, B200MPPI, u_tol=1e-4, cost_rtol=1e-5)
            barrier()
        if rank == 0:
            parity = dict(p64)
            # the run fails on: the float64 check (merge algebra exact to 1e-9), ranks that disagree bitwise, or per-sample
            # costs outside the 1e-5 gate in the bench precision; the float32-mode U distance is reported (`ok` there)
            parity["ok"] = bool(p64["ok"] and (pb is None or (pb["U_bitwise_equal_across_ranks"] and pb["cost_max_rel_vs_unsharded"] < 1e-5)))
            if pb is not None:
                parity["bench_precision"] = pb

    mppi = B200MPPI(spec, precision=args.precision, device=local_rank, rank=rank, world_size=world, seed=1234,
                    U_init=torch.from_numpy(np.asarray(z['in.U']).copy()), use_torch_stream=True, max_horizon=max(T, 64))
    stream = mppi.torch_stream
    kernel_name = mppi.kernel_name

    # ---------------- device-timed throughput (state resident, L2 flushed between steps)
    mppi.set_state_resident(state)

    def device_step():
        if world == 1 or mppi._p2p:
            mppi.command_resident()  # sharded: the merge kernel exchanges the partials over peer memory itself
        else:
            mppi.command(state)      # NCCL fallback: the all-gather is issued from the host

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
    p2p = bool(mppi._p2p)
    mppi.close()

    # ---------------- the other records of the line (outside the timed regions above)
    strong_pts = None
    if not args.no_extras and not strong and args.workload == 'push':
        strong_pts = strong_scaling_points(args, rank, local_rank, world, dist, torch, B200MPPI, flush, barrier)
    f64_rec = None
    if not args.no_extras and world == 1 and args.precision != 'f64' and spec.dynamics.kind != 2:
        m64 = B200MPPI(spec, precision='f64', device=local_rank, seed=1234, U_init=torch.from_numpy(np.asarray(z['in.U']).copy()),
                       use_torch_stream=True, max_horizon=max(T, 64))
        ms64 = float(timed_commands(m64, state, min(args.steps, 20), 3, flush, barrier, torch).mean())
        for _ in range(3):
            m64.rollout_resident()
        torch.cuda.synchronize()
        evr = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for e0, e1 in evr:
            with torch.cuda.stream(m64.torch_stream):
                flush.zero_()
                e0.record()
                m64.rollout_resident()
                e1.record()
        torch.cuda.synchronize()
        r64 = float(np.mean([e0.elapsed_time(e1) for e0, e1 in evr]))
        tf = K * T * 2.0 * spec.dynamics.macs_per_step / (r64 * 1e-3) * 1e-12
        f64_rec = {"value": K * T / (ms64 * 1e-3), "unit": UNIT, "ms_per_step": ms64, "kernel_ms": r64, "achieved_tflops": tf,
                   "peak_tflops": FP64_TENSOR_PEAK_TFLOPS, "frac": tf / FP64_TENSOR_PEAK_TFLOPS, "kernel": m64.kernel_name,
                   "note": "the reference's own arithmetic (controller.py:197 torch.double) on the FP64 tensor pipe; parity 1e-9"}
        m64.close()

    if rank == 0:
        value = K * T / (ms_per_step * 1e-3)
        e2e_val = K * T * args.steps / e2e_wall
        flop_per_step = 2.0 * spec.dynamics.macs_per_step  # 7046 for push
        roofline = None
        fma_kernel = 'rollout_kernel' in kernel_name and 'mma' not in kernel_name
        if roll_ms is not None:
            achieved = K * T * flop_per_step / (roll_ms * 1e-3) * 1e-12
            # the kernel is compute-bound (it moves ~100 KB of DRAM traffic per launch); which pipe bounds it
            # depends on the arithmetic THAT RAN (kernel_name): every peak below was measured on this pool's B200 with
            # tools/peaks (MEASURED_PEAKS.json holds only the HBM copy and the cuBLAS bf16 tcgen05 GEMM)
            if fma_kernel or spec.dynamics.kind == 2:
                bound, peak = "fp32_fma", FP32_PEAK_TFLOPS
                src = "FFMA2 register-operand peak measured (profiles/fma_peaks_r1.jsonl)"
            elif 'tcgen05' in kernel_name:
                bound, peak = "tensor", TCGEN05_TF32_PEAK_TFLOPS / 3.0
                src = ("tcgen05.mma kind::tf32 M=128 measured {:.0f} TFLOP/s (profiles/tcgen05_probe_r2.jsonl) / 3 products per "
                       "algorithmic MAC (error-compensated 3xTF32 split)").format(TCGEN05_TF32_PEAK_TFLOPS)
            elif '3xTF32' in kernel_name:
                bound, peak = "tensor", TF32_MMA_PEAK_TFLOPS / 3.0
                src = ("mma.sync m16n8k8 tf32 measured {:.0f} TFLOP/s (profiles/mma_peaks_r1.jsonl) / 3 products per "
                       "algorithmic MAC (error-compensated 3xTF32 split)").format(TF32_MMA_PEAK_TFLOPS)
            elif '3xFP16' in kernel_name:
                bound, peak = "tensor", F16_MMA_PEAK_TFLOPS / 3.0
                src = ("mma.sync m16n8k16 f16 measured {:.0f} TFLOP/s (profiles/mma_peaks_r1.jsonl) / 3 products per "
                       "algorithmic MAC (error-compensated 3xFP16 split)").format(F16_MMA_PEAK_TFLOPS)
            elif 'DMMA' in kernel_name:
                bound, peak = "tensor", FP64_TENSOR_PEAK_TFLOPS
                src = "mma.sync m8n8k4 f64 (DMMA) measured (profiles/mma_peaks_r1.jsonl)"
            else:
                bound, peak = "fp32_fma", FP32_PEAK_TFLOPS
                src = "FFMA2 register-operand peak measured (profiles/fma_peaks_r1.jsonl)"
            roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                        "frac": achieved / peak,
                        "frac_fp32_fma": achieved / FP32_PEAK_TFLOPS,
                        "frac_tcgen05_tf32x3": achieved / (TCGEN05_TF32_PEAK_TFLOPS / 3.0) if TCGEN05_TF32_PEAK_TFLOPS else None,
                        "traffic": NCU_DRAM_BYTES_PER_LAUNCH.get(args.precision) if args.workload == 'push' else None,
                        "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size, one "
                                          "ncu --set full capture (profiles/r2z_split3_k8192_ncu_summary.txt; f16x3: r1e_rollout_split_f16x3_k8192_ncu_summary.txt)",
                        "kernel": "rollout (fused K x T): " + kernel_name,
                        "kernel_ms": roll_ms, "algorithmic_flop_per_launch": K * T * flop_per_step,
                        "peak_source": src,
                        "frac_note": "frac is against the pipe the kernel ran on; frac_fp32_fma against the measured FFMA2 peak "
                                     "(SURVEY.md §8d's stated bound); frac_tcgen05_tf32x3 against the measured tcgen05 kind::tf32 peak / 3"}
        # CPU baseline: bounded sample of the same workload on the host cores (oracle port, float64)
        torch.set_num_threads(os.cpu_count() or 1)
        Kc = cpu_sample_K(args.workload)
        zc, spec_c = load_spec(args.workload, 1, Kc)
        Uc = torch.from_numpy(np.asarray(z['in.U']).copy())
        cpu_reference_step(spec_c, zc, state, Uc, 1)
        cpu_t = [cpu_reference_step(spec_c, zc, state, Uc, 2 + i) for i in range(args.cpu_steps)]
        cpu_val = Kc * T / float(np.mean(cpu_t))
        cuda_t = torch_cuda_f64_comparator(spec_c, zc, state, Uc) if world == 1 else None
        m10 = cpu_reference_m10(args.workload, state, Uc) if (world == 1 and not args.no_extras and spec.dynamics.kind != 2) else None
        if fma_kernel or spec.dynamics.kind == 2:
            dtype = "f32 arithmetic (FFMA2), f64 state+cost" if args.precision in ('mixed', 'tf32x3', 'f16x3') else args.precision
        else:
            dtype = {"f64": "f64", "mixed": "f32 networks, f64 state+cost", "f32": "f32",
                     "tf32x3": "3xTF32 (f32-equivalent) networks, f32 state, f64 cost sum",
                     "f16x3": "3xFP16 (f32-equivalent) networks, f32 state, f64 cost sum"}.get(args.precision, args.precision)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": make_config(args.workload, world, K, T),
            "precision": args.precision, "kernel": kernel_name,
            "collective": None if world == 1 else ("p2p-ipc (one ~1 KB partial per rank stored into CUDA-IPC mapped peer memory by the merge "
                                                   "kernel; NCCL is the bootstrap and the timing all-reduce only)" if p2p else
                                                   "nccl all_gather_into_tensor of one tampc_partial per rank"),
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
        if parity is not None:
            line["shard_parity"] = parity
        if strong_pts is not None:
            line["strong_K1M"] = strong_pts
        if f64_rec is not None:
            line["f64"] = f64_rec
        if m10 is not None:
            line["reference_M10"] = m10
        print(json.dumps(line), flush=True)
    rc = 0
    if world > 1:
        ok = torch.tensor([1 if (parity is None or parity.get("ok")) else 0], dtype=torch.int32, device='cuda')
        dist.broadcast(ok, src=0)
        rc = 0 if int(ok[0]) else 1
        dist.destroy_process_group()
    if rc:
        sys.stderr.write("shard parity FAILED: the sharded command does not reproduce the unsharded one\n")
        sys.exit(rc)


if __name__ == '__main__':
    main()
