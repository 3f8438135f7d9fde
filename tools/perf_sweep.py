"""Quick timing sweep of the CUDA path (development tool; bench.py is the contract harness).

  python tools/perf_sweep.py [--case push_nominal] [--precision mixed] [--K 8192 65536 1048576]
Prints one JSON line per (K, samples/thread, block) with rollout-kernel, full-command and end-to-end times."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tampc_b200 import spec as S  # noqa: E402
from tampc_b200.mppi import B200MPPI  # noqa: E402


def time_events(fn, iters, stream, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case', default='push_nominal')
    ap.add_argument('--precision', nargs='+', default=['mixed'])
    ap.add_argument('--K', type=int, nargs='+', default=[8192, 65536, 1048576])
    ap.add_argument('--T', type=int, default=0)
    ap.add_argument('--spt', type=int, nargs='+', default=[0])
    ap.add_argument('--block', type=int, nargs='+', default=[0])
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--synthetic', type=int, default=0, help='hidden width of the synthetic MLP problem (config 4)')
    args = ap.parse_args()
    z = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', args.case + '.npz'))
    for prec in args.precision:
        for K in args.K:
            for spt in args.spt:
                for block in args.block:
                    if args.synthetic:
                        from tampc_b200.synthetic import make_synthetic_spec
                        spec, x0 = make_synthetic_spec(H=args.synthetic, K=K, T=args.T or 40)
                        z = {'in.U': np.zeros((spec.sampler.T, 3)), 'in.state': x0}
                    else:
                        spec = S.spec_from_arrays(z)
                        spec.sampler.K = K
                        if args.T:
                            spec.sampler.T = args.T
                    T = spec.sampler.T
                    if spt:
                        os.environ['TAMPC_SAMPLES_PER_THREAD'] = str(spt)
                    if block:
                        os.environ['TAMPC_BLOCK'] = str(block)
                    m = B200MPPI(spec, precision=prec, use_torch_stream=True, max_horizon=max(T, 64),
                                 U_init=torch.from_numpy(np.resize(z['in.U'], (T, spec.sampler.nu))))
                    m.set_state_resident(z['in.state'])
                    iters = max(3, min(args.iters, int(2e9 / (K * T * 3600)) + 3))
                    t_roll = time_events(m.rollout_resident, iters, m.torch_stream)
                    t_cmd = time_events(m.command_resident, iters, m.torch_stream)
                    state = z['in.state']
                    for _ in range(3):
                        m.command(state)
                    t0 = time.perf_counter()
                    for _ in range(iters):
                        m.command(state)
                    t_e2e = (time.perf_counter() - t0) / iters * 1e3
                    flops = 2.0 * spec.dynamics.macs_per_step * K * T
                    print(json.dumps({'case': ('synthetic_H%d' % args.synthetic) if args.synthetic else args.case, 'precision': prec, 'K': K, 'T': T, 'spt': spt, 'block': block,
                                      'rollout_ms': round(t_roll, 4), 'command_ms': round(t_cmd, 4), 'e2e_ms': round(t_e2e, 4),
                                      'rollout_tflops': round(flops / t_roll * 1e-9, 2),
                                      'Msteps_per_s_e2e': round(K * T / t_e2e * 1e-3, 1)}), flush=True)
                    m.close()


if __name__ == '__main__':
    main()
