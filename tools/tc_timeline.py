"""Development tool: where one horizon step of the tcgen05 MLP kernel goes (SM-clock stamps of CTA 0's roles at step 2).
   TAMPC_TC_PROF=1 python tools/tc_timeline.py --H 256 --K 1024"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault('TAMPC_TC_PROF', '1')
from tampc_b200.mppi import B200MPPI  # noqa: E402
from tampc_b200.synthetic import make_synthetic_spec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--H', type=int, default=256)
ap.add_argument('--K', type=int, nargs='+', default=[1024, 1048576])
args = ap.parse_args()
for K in args.K:
    spec, x0 = make_synthetic_spec(H=args.H, K=K, T=40)
    m = B200MPPI(spec, precision='tf32x3', U_init=torch.zeros(40, 3, dtype=torch.double))
    for _ in range(3):
        m.command(x0)
    out = (C.c_int64 * 64)()
    m._check(m._lib.tampc_mppi_debug_stamps(m._h, out))
    s = np.array(out[:], dtype=np.int64)
    t0 = s[0]
    rel = lambda i: int(s[i] - t0) if s[i] else None
    print('H', args.H, 'K', K, m.kernel_name[:40])
    print('  owners   : step start 0 | A_in arrived', rel(1), '| D1(0) full', rel(2), '| A1(0) written', rel(3), '| last E1 done', rel(4),
          '| presampled', rel(5), '| D2 full', rel(6), '| E2 done', rel(7), '| exchanged', rel(8), '| x updated', rel(9))
    print('  group 1  : D1(1) full', rel(48), '| A1(1) written', rel(49), '| D2 full', rel(50), '| E2 done', rel(51))
    print('  L1 issue : woke', rel(32), '| chunks committed', [rel(33 + c) for c in range(8)])
    print('  L2 issue : woke', rel(16), '| chunks committed', [rel(17 + c) for c in range(8)], '| D2 commit', rel(26))
    m.close()
