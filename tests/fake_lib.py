"""TEST INFRASTRUCTURE: a stand-in for libtampc_mppi.so on a box without a GPU.

It speaks the part of the C ABI (include/tampc_mppi.h) that B200MPPI uses through ctypes — create / set_model / set_gp /
set_cost / set_nominal / get_nominal / command / predict / predict_nominal / state_costs / query — records every
descriptor that crosses the boundary, and answers `command` with the CPU oracle run on the spec object of the bound
B200MPPI.  With it the JOIN between the real reference controller and the drop-in (keyword constructor, deferred cost
extraction, per-command refresh, attribute surface, tensors in and out) runs on the CPU box; the kernels behind the same
calls are checked on the GPU box against the same oracle."""
import copy
import ctypes as C

import numpy as np
import torch

from oracle import mppi_oracle as O
from tampc_b200 import _lib as L
from tampc_b200 import spec as S


def _arr(ptr, n):
    return np.ctypeslib.as_array(ptr, shape=(n,))


def _clone(struct):
    """A by-value copy of a ctypes structure (what crossing the C ABI by pointer-to-const amounts to)."""
    out = type(struct)()
    C.memmove(C.byref(out), C.byref(struct), C.sizeof(struct))
    return out


class FakeLib:
    def __init__(self):
        self.calls = []
        self.mppi = None
        self.U = None
        self.T = 0
        self.next_noise = None   # (K, T, nu) draw N(mu, Sigma) for the next command, like patching noise_dist.sample
        self.last = None
        self.launches = 0
        self.model_desc = self.cost_desc = self.cfg = None

    def bind(self, mppi):
        self.mppi = mppi

    # ---- plumbing
    def tampc_mppi_last_error(self, h):
        return b"fake"

    def tampc_mppi_abi_version(self):
        return L.ABI_VERSION

    def tampc_mppi_kernel_name(self, h):
        return b"cpu oracle behind the C ABI (tests/fake_lib.py)"

    def tampc_mppi_create(self, cfg, out):
        self.cfg = _clone(cfg._obj)
        out._obj.value = 1
        self.calls.append('create')
        return L.OK

    def tampc_mppi_destroy(self, h):
        self.calls.append('destroy')

    def tampc_mppi_set_model(self, h, desc, params, n):
        self.model_desc = _clone(desc._obj)
        self.model_params = None if not n else _arr(params, n).copy()
        self.calls.append('set_model')
        return L.OK

    def tampc_mppi_set_gp(self, h, desc, params, n):
        self.gp_desc = _clone(desc._obj)
        self.calls.append('set_gp')
        return L.OK

    def tampc_mppi_set_cost(self, h, desc):
        self.cost_desc = _clone(desc._obj)
        self.calls.append('set_cost')
        return L.OK

    def tampc_mppi_set_nominal(self, h, U, T):
        nu = self.cfg.nu
        self.U = _arr(U, T * nu).reshape(T, nu).copy()
        self.T = T
        return L.OK

    def tampc_mppi_get_nominal(self, h, out):
        _arr(out, self.U.size)[:] = self.U.reshape(-1)
        return L.OK

    # ---- the path
    def _spec(self):
        m = self.mppi
        spec = S.RolloutSpec(m._dyn_spec, m._cost_spec, copy.copy(m.spec.sampler))
        spec.sampler.T = self.T
        return spec

    def tampc_mppi_command(self, h, state, mode, noise, action):
        spec = self._spec()
        nx, nu, K, T = self.cfg.nx, self.cfg.nu, self.cfg.num_samples, self.T
        x = torch.from_numpy(_arr(state, nx).copy())
        if mode == L.NOISE_INJECTED:
            nz = torch.from_numpy(_arr(noise, K * T * nu).reshape(K, T, nu).copy())
        else:
            assert self.next_noise is not None, "the test must provide the draw of this command"
            nz, self.next_noise = self.next_noise, None
        out = O.command(spec, x, torch.from_numpy(self.U), nz)
        self.U = out['U'].numpy().copy()
        self.last = out
        _arr(action, nu)[:] = out['action'].numpy()
        self.launches += 2
        self.calls.append('command')
        return L.OK

    def _predict(self, state, act, out, nominal):
        d = self.mppi._dyn_spec
        x = torch.from_numpy(_arr(state, d.nx).copy()).view(1, -1)
        u = torch.from_numpy(_arr(act, d.nu).copy()).view(1, -1)
        if nominal:
            dn = copy.copy(d)
            dn.gp = None
            nxt = O.predict(dn, x, u)
        else:
            nxt, _ = O.apply_dynamics(d, x, u, None if d.gp is None else torch.zeros(1, d.gp.resid.shape[1], dtype=torch.double))
        _arr(out, d.nx)[:] = nxt[0].numpy()
        self.calls.append('predict_nominal' if nominal else 'predict')
        return L.OK

    def tampc_mppi_predict(self, h, state, act, out):
        return self._predict(state, act, out, False)

    def tampc_mppi_predict_nominal(self, h, state, act, out):
        return self._predict(state, act, out, True)

    def tampc_mppi_get_rollouts(self, h, state, out):
        spec = self._spec()
        x = torch.from_numpy(_arr(state, self.cfg.nx).copy())
        r = O.get_rollouts(spec, x, torch.from_numpy(self.U))
        _arr(out, r.numel())[:] = r.reshape(-1).numpy()
        return L.OK

    def tampc_mppi_state_costs(self, h, states, n, goal_out, trap_out):
        c = self.mppi._cost_spec
        xs = torch.from_numpy(_arr(states, n * c.nx).reshape(n, c.nx).copy())
        for i in range(n):
            _, g, t = O.normalize_trapset_cost_to_state(c, xs[i])
            _arr(goal_out, n)[i] = float(g)
            _arr(trap_out, n)[i] = float(t) / c.trap_weight
        self.calls.append('state_costs')
        return L.OK

    def tampc_mppi_query(self, h, what, out, n):
        if what == L.Q_STATS:
            _arr(out, 4)[:] = [float(self.last['beta']), float(self.last['eta']), self.last['argmin'], self.launches]
        elif what == L.Q_COST_TOTAL:
            _arr(out, n)[:] = self.last['cost_total'].numpy()
        elif what == L.Q_PERTURBED:
            _arr(out, n)[:] = self.last['perturbed_action'].reshape(-1).numpy()
        elif what == L.Q_NOISE:
            _arr(out, n)[:] = self.last['noise'].reshape(-1).numpy()
        elif what == L.Q_OMEGA:
            _arr(out, n)[:] = self.last['omega'].numpy()
        else:
            return L.ERR_INVALID
        return L.OK
