"""Rollout kernel, whole command and their difference (the softmax-update tail) at the bench's K=8192, with and without
the L2 flush between commands:  python tools/flush_test.py   (B200; device-timed with CUDA events)."""
import sys, os, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from conftest import load_golden
from tampc_b200.mppi import B200MPPI
z, spec = load_golden('push_nominal'); spec.sampler.K = 8192
m = B200MPPI(spec, precision='tf32x3', seed=1, U_init=torch.from_numpy(np.asarray(z['in.U']).copy()), use_torch_stream=True, max_horizon=64)
m.set_state_resident(z['in.state'].copy())
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def run(fn, do_flush, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for e0, e1 in ev:
        with torch.cuda.stream(m.torch_stream):
            if do_flush: flush.zero_()
            e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev])) * 1e3
for fl in (1, 0):
    c = run(m.command_resident, fl); r = run(m.rollout_resident, fl)
    print({"flush": fl, "command_us": round(c, 2), "rollout_us": round(r, 2), "tail_us": round(c - r, 2)})
