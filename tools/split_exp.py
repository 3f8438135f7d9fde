"""Experiment driver: rollout time of the push model at latency-bound K plus a checksum of the costs (same seed), for
whatever TAMPC_* switches are set in the environment."""
import hashlib
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from tampc_b200 import spec as S  # noqa: E402
from tampc_b200.mppi import B200MPPI  # noqa: E402


def main():
    z = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'push_nominal.npz'))
    out = {'env': {k: v for k, v in os.environ.items() if k.startswith('TAMPC_')}}
    for K in [int(k) for k in sys.argv[1:]] or [4736, 8192, 12288]:
        spec = S.spec_from_arrays(z)
        spec.sampler.K = K
        T = spec.sampler.T
        m = B200MPPI(spec, precision='tf32x3', use_torch_stream=True, max_horizon=64, seed=11,
                     U_init=torch.from_numpy(np.resize(z['in.U'], (T, spec.sampler.nu))))
        m.set_state_resident(z['in.state'])
        m.command(z['in.state'])
        h = hashlib.sha1(m.cost_total.numpy().tobytes()).hexdigest()[:10]
        best = []
        for fn in (m.rollout_resident, m.command_resident):
            ts = []
            for _ in range(5):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(m.torch_stream):
                    e0.record()
                    for _ in range(20):
                        fn()
                    e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) / 20)
            best.append(round(float(np.median(ts)) * 1e3, 2))
        out[str(K)] = {'rollout_us': best[0], 'command_us': best[1], 'cost_sha': h, 'kernel': m.kernel_name[:28]}
        m.close()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
