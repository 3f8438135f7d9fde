"""Timing of the device GP fit (tampc_gp_fit) on the recorded stages of tests/golden/gp_fit.npz."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tampc_b200 import _lib as L  # noqa: E402

lib = L.load()
z = np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'gp_fit.npz'))
for N in (34, 50):
    X = np.ascontiguousarray(np.resize(z['stage4.X'], (N, 8)) + 0.01 * np.arange(N)[:, None])
    R = np.ascontiguousarray(np.resize(z['stage4.resid'], (N, 5)))
    for iters in (15, 150):
        ts = []
        for rep in range(6):
            raw, m, v, step, loss = np.zeros(16), np.zeros(16), np.zeros(16), C.c_int64(0), np.zeros(2)
            t0 = time.perf_counter()
            rc = lib.tampc_gp_fit(0, N, 8, 5, L.dptr(X), L.dptr(R), iters, 0.1, L.dptr(raw), L.dptr(m), L.dptr(v), C.byref(step), L.dptr(loss))
            ts.append(time.perf_counter() - t0)
            assert rc == 0
        print('{{"bench": "gp_fit", "n_points": {}, "iters": {}, "ms_per_call": {:.3f}}}'.format(N, iters, 1e3 * float(np.median(ts[1:]))))
