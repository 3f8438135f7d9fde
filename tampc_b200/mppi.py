"""B200MPPI — drop-in for the `pytorch_mppi.MPPI` / `ExperimentalMPPI` object that tampc's controllers hold as
`ctrl.mpc` (tampc/controller/controller.py:316-435; plug point :418-421; methods and attributes touched from
outside listed in SURVEY.md §8b).

The object keeps the reference's surface — `command(state)`, `reset()`, `get_rollouts(state)`,
`change_horizon(T)`, `_compute_rollout_costs(actions)`, attributes `K, T, U, nx, nu, M, lambda_, noise_*,
running_cost, terminal_state_cost, cost_total, omega, perturbed_action, state, d, dtype` — but the rollout and the
softmax update run in the hand-written sm_100a kernels behind the C ABI of include/tampc_mppi.h.  Tensors at the
API are float64 torch tensors on the host, like the reference's `self.dtype = torch.double`
(controller.py:197); conversion happens at the edge.

There is no CPU path: constructing a B200MPPI without the CUDA library or a B200 raises.
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from . import extract
from . import spec as S


class TampcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libtampc_mppi error {}: {}".format(code, message))
        self.code = code


def shard_bounds(K, rank, world):
    """Contiguous K-shard of `rank`: sizes differ by at most one; offsets are a function of (K, world) only."""
    base, rem = divmod(K, world)
    lo = rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


class B200MPPI:
    def __init__(self, *args, **kwargs):
        """Two forms.

        B200MPPI(spec: RolloutSpec, precision='mixed', device=0, ...)  — from a flattened problem (see `_init_from_spec`).

        B200MPPI(dynamics, running_cost, nx, noise_sigma, num_samples=100, horizon=15, device='cpu',
                 terminal_state_cost=None, lambda_=1., noise_mu=None, u_min=None, u_max=None, u_init=None, U_init=None,
                 u_scale=1, u_per_command=1, step_dependent_dynamics=False, sample_null_action=False, rollout_samples=20,
                 rollout_var_cost=0.5, rollout_var_discount=0.95, **b200_options)
        — the `pytorch_mppi.MPPI` / `ExperimentalMPPI` constructor (controller.py:316-321, the call at :418-421), so that the
        reference's own `self.mpc = ExperimentalMPPI(self._apply_dynamics, self._running_cost, self.nx, ...)` line can build
        this class (`b200_controller_class`).  `dynamics` / `running_cost` / `terminal_state_cost` must be bound methods of a
        tampc controller: the kernels cannot call Python, the objects behind the callbacks are read instead (extract.py)."""
        if (args and isinstance(args[0], S.RolloutSpec)) or 'spec' in kwargs:
            self._init_from_spec(*args, **kwargs)
        else:
            self._init_like_pytorch_mppi(*args, **kwargs)

    def _init_like_pytorch_mppi(self, dynamics, running_cost, nx, noise_sigma, num_samples=100, horizon=15, device='cpu',
                                terminal_state_cost=None, lambda_=1., noise_mu=None, u_min=None, u_max=None, u_init=None,
                                U_init=None, u_scale=1, u_per_command=1, step_dependent_dynamics=False, sample_null_action=False,
                                rollout_samples=20, rollout_var_cost=0.5, rollout_var_discount=0.95, cuda_device=0, **b200):
        ctrl = getattr(dynamics, '__self__', None)
        if ctrl is None or not hasattr(ctrl, 'dynamics'):
            raise extract.UnsupportedSpec("`dynamics` must be a bound method of a tampc controller (e.g. ctrl._apply_dynamics): the "
                                          "fused kernels read the model behind the callback, they cannot call Python")
        if u_scale != 1 or u_per_command != 1:
            raise extract.UnsupportedSpec("u_scale / u_per_command other than 1 are not used by the reference controllers")
        if int(nx) != int(ctrl.nx):
            raise ValueError("nx={} differs from the controller's {}".format(nx, ctrl.nx))
        ns = extract._np(noise_sigma)
        nu = 1 if ns.ndim == 0 else ns.shape[0]
        ns = ns.reshape(nu, nu)
        mu = np.zeros(nu) if noise_mu is None else extract._np(noise_mu).reshape(nu)
        lo = None if u_min is None else np.broadcast_to(extract._np(u_min), (nu,)).copy()
        hi = None if u_max is None else np.broadcast_to(extract._np(u_max), (nu,)).copy()
        if hi is not None and lo is None:  # pytorch_mppi mirrors a missing bound
            lo = -hi
        if lo is not None and hi is None:
            hi = -lo
        sampler = S.SamplerSpec(K=int(num_samples), T=int(horizon), nu=nu, lambda_=float(lambda_), noise_mu=mu, noise_sigma=ns,
                                u_min=lo, u_max=hi, u_init=None if u_init is None else extract._np(u_init).reshape(nu),
                                sample_null_action=bool(sample_null_action), rollout_samples=int(rollout_samples),
                                rollout_var_cost=float(rollout_var_cost), rollout_var_discount=float(rollout_var_discount))
        dyn = extract.dynamics_from_controller(ctrl)
        # the cost program is read on the first command(): MPPI_MPC builds its mpc before set_goal() can have been called
        placeholder = S.CostSpec(kind=S.COST_GOAL_TRAP, nx=dyn.nx, nu=nu, Q=np.zeros((dyn.nx, dyn.nx)), R=np.zeros((nu, nu)),
                                 goal=np.zeros(dyn.nx), use_trap_cost=False)
        b200.setdefault('precision', 'tf32x3')
        self._init_from_spec(S.RolloutSpec(dyn, placeholder, sampler, name=type(ctrl).__name__), device=cuda_device, controller=ctrl,
                             U_init=None if U_init is None else torch.as_tensor(U_init).detach().to('cpu', torch.double), **b200)
        self.running_cost = running_cost
        self.terminal_state_cost = terminal_state_cost
        self.F = dynamics
        self.step_dependency = bool(step_dependent_dynamics)
        self.d = torch.device(device) if not isinstance(device, torch.device) else device  # where the API tensors live

    def _init_from_spec(self, spec: S.RolloutSpec, precision='mixed', device=0, seed=1234, max_horizon=None,
                        rank=0, world_size=1, process_group=None, controller=None, U_init=None, use_torch_stream=False,
                        exchange='auto'):
        """
        :param spec: flattened problem (dynamics, cost, sampler constants)
        :param precision: 'f64' (reference arithmetic), 'mixed' (float32 networks, float64 state/cost), 'f32'
        :param rank, world_size, process_group: K-sharding over torch.distributed ranks (one process per GPU);
               every rank holds the full U and obtains bitwise-identical updates
        :param controller: live tampc controller to re-read the cost program from on every command()
        :param exchange: multi-GPU exchange of the softmax partials: 'p2p' = stores into CUDA-IPC mapped peer memory
               from inside the merge kernel (one node, NVLink); 'nccl' = all_gather_into_tensor; 'auto' = p2p, falling
               back to nccl if the peers cannot be mapped
        """
        self._lib = L.load()
        self.spec = spec
        self.precision = precision
        self.device_index = device
        self.rank, self.world_size, self.group = rank, world_size, process_group
        self._ctrl = controller
        s = spec.sampler
        self.K, self.T = s.K, s.T
        self.nx, self.nu = spec.dynamics.nx, s.nu
        self.M = s.rollout_samples
        self.rollout_var_cost, self.rollout_var_discount = s.rollout_var_cost, s.rollout_var_discount
        self.lambda_ = s.lambda_
        self.dtype = torch.double
        self.d = torch.device('cpu')  # tensors at the API live on the host, like MPC.command's numpy edge
        self.noise_mu = torch.from_numpy(s.noise_mu.copy())
        self.noise_sigma = torch.from_numpy(s.noise_sigma.copy())
        self.noise_sigma_inv = torch.from_numpy(s.noise_sigma_inv.copy())
        self.noise_dist = torch.distributions.MultivariateNormal(self.noise_mu, covariance_matrix=self.noise_sigma)
        self.u_min = None if s.u_min is None else torch.from_numpy(s.u_min.copy())
        self.u_max = None if s.u_max is None else torch.from_numpy(s.u_max.copy())
        self.u_init = torch.from_numpy(s.u_init.copy())
        self.sample_null_action = s.sample_null_action
        self.u_per_command = 1
        self.running_cost = None
        self.terminal_state_cost = None
        self._state_np = None
        self._x_buf, self._act_buf = np.zeros(self.nx), np.zeros(self.nu)  # command()'s persistent host buffers
        self._x_ptr, self._act_ptr = L.dptr(self._x_buf), L.dptr(self._act_buf)
        self.step_dependency = True

        self.sample_offset, self.K_local = shard_bounds(self.K, rank, world_size)
        if self.K_local < 1:
            raise ValueError("rank {} of {} has no samples (K={})".format(rank, world_size, self.K))
        stream = None
        self.torch_stream = None
        if world_size > 1 or use_torch_stream:
            # a torch-owned side stream shared with the library: NCCL collectives and torch.cuda.Event timing issued
            # under `with torch.cuda.stream(self.torch_stream)` are ordered with the kernels (the legacy default
            # stream has handle 0, which the C ABI reads as "create your own")
            torch.cuda.set_device(device)
            self.torch_stream = torch.cuda.Stream(device=device)
            stream = C.c_void_p(self.torch_stream.cuda_stream)
        cfg = L.mppi_config(s, self.nx, precision, device=device, sample_offset=self.sample_offset,
                            num_local=self.K_local, max_horizon=max_horizon or max(s.T, 64), seed=seed, stream=stream)
        self.max_horizon = cfg.max_horizon
        h = C.c_void_p()
        rc = self._lib.tampc_mppi_create(C.byref(cfg), C.byref(h))
        if rc != L.OK:
            raise TampcError(rc, self._lib.tampc_mppi_last_error(None).decode())
        self._h = h
        self._cfg = cfg
        self._cost_spec = None
        self._dyn_spec = None
        self.set_dynamics(spec.dynamics)
        self.set_cost(spec.cost)
        if U_init is None:
            U_init = self.noise_dist.sample((self.T,))
        self.U = U_init
        self._gather = None
        self._p2p = False
        if world_size > 1 and exchange in ('auto', 'p2p'):
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self._connect_peers(required=exchange == 'p2p')
            elif exchange == 'p2p':
                raise RuntimeError("exchange='p2p' needs an initialised torch.distributed process group to trade the IPC handles")
            # 'auto' without a process group: the caller drives command_begin / command_finish with its own transport
        self._last_noise = None
        self._cost_fp = None
        self._cost_cache = None
        self._probe_cache = {}

    # ------------------------------------------------------------------ construction from a live controller
    @classmethod
    def from_controller(cls, ctrl, **kwargs):
        """Build from a tampc MPPI_MPC / OnlineMPPI controller (reads ctrl.mpc for the sampler constants)."""
        spec = extract.spec_from_controller(ctrl)
        old = ctrl.mpc
        obj = cls(spec, controller=ctrl, U_init=old.U.detach().to('cpu', torch.double), **kwargs)
        obj.running_cost = old.running_cost
        obj.terminal_state_cost = old.terminal_state_cost
        return obj

    @classmethod
    def from_package(cls, path, **kwargs):
        """Build from a packaged spec written by `tampc_b200.spec.save_spec`: no datasets, no scaler
        fitting and none of the reference's un-vendored libraries are needed at start-up."""
        return cls(S.load_spec(path), **kwargs)

    @classmethod
    def for_one_step(cls, ctrl, samples, this_cost='none', **kwargs):
        """A handle for the one-step action samplers of the APF baselines (APFVO / APFSP, online_controller.py:554-664)
        and anything else that pushes `samples` candidate actions through `ctrl._apply_dynamics` from one state: the
        dynamics of `ctrl`, K = samples, horizon 1.

        this_cost='none': dynamics only (use `apply_dynamics`).  this_cost=None / a cost object of the controller: also the
        APF potential (extract.apf_cost_from_controller), so that `select_action` runs dynamics + cost + argmin on the
        device; `set_apf_cost` re-reads it (APFVO's trap set grows, APFSP switches between goal and helicoid potential)."""
        dyn = extract.dynamics_from_controller(ctrl)
        nx, nu = dyn.nx, dyn.nu
        if isinstance(this_cost, str) and this_cost == 'none':
            cost = S.CostSpec(kind=S.COST_GOAL_TRAP, nx=nx, nu=nu, Q=np.zeros((nx, nx)), R=np.zeros((nu, nu)), goal=np.zeros(nx),
                              use_trap_cost=False)
        else:
            cost = extract.apf_cost_from_controller(ctrl, this_cost)
        u_min = None if getattr(ctrl, 'u_min', None) is None else extract._np(ctrl.u_min)
        u_max = None if getattr(ctrl, 'u_max', None) is None else extract._np(ctrl.u_max)
        sampler = S.SamplerSpec(K=int(samples), T=1, nu=nu, lambda_=1.0, noise_mu=np.zeros(nu), noise_sigma=np.eye(nu),
                                u_min=u_min, u_max=u_max)
        obj = cls(S.RolloutSpec(dyn, cost, sampler, name='one_step'), U_init=torch.zeros(1, nu, dtype=torch.double), **kwargs)
        obj._apf_ctrl = ctrl
        return obj

    def set_apf_cost(self, this_cost=None):
        """Re-read the APF potential from the controller this one-step handle was built for."""
        self.set_cost(extract.apf_cost_from_controller(self._apf_ctrl, this_cost))

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != L.OK:
            raise TampcError(rc, self._lib.tampc_mppi_last_error(self._h).decode())

    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.tampc_mppi_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_dynamics(self, dyn: S.DynamicsSpec):
        desc, params = L.model_desc(dyn)
        self._check(self._lib.tampc_mppi_set_model(self._h, C.byref(desc), L.dptr(params) if params.size else None,
                                                   params.size))
        self._dyn_spec = dyn
        self.spec.dynamics = dyn
        self.set_gp(dyn.gp)

    def set_gp(self, gp: Optional[S.GPSpec]):
        """Online GP residual around the nominal network (OnlineGPMixing, online_model.py:324-457), frozen at its
        current data / hyper-parameters; None switches it off.  Cheap (<= 64 points): call after every GP update."""
        s = self.spec.sampler
        if gp is not None and not 1 <= self.M <= L.MAX_ROLLOUT_SAMPLES:
            raise extract.UnsupportedSpec("rollout_samples={} outside [1,{}]".format(self.M, L.MAX_ROLLOUT_SAMPLES))
        s.rollout_samples, s.rollout_var_cost, s.rollout_var_discount = self.M, self.rollout_var_cost, self.rollout_var_discount
        desc, params = L.gp_desc(gp, s)
        self._check(self._lib.tampc_mppi_set_gp(self._h, C.byref(desc), L.dptr(params) if params.size else None, params.size))
        self._dyn_spec.gp = gp

    def set_cost(self, cost: S.CostSpec):
        desc = L.cost_desc(cost)
        self._check(self._lib.tampc_mppi_set_cost(self._h, C.byref(desc)))
        self._cost_spec = cost
        self.spec.cost = cost

    def _refresh_from_controller(self):
        """Re-read the mutable parts of the problem from the bound controller: cost program (trap set, trap weight,
        goal, recovery sets / MAB weights: online_controller.py:353,400-402,416-418,461-478)."""
        if self._ctrl is None:
            return
        fp = self._cost_fingerprint()
        if fp != self._cost_fp:
            cost = extract.cost_from_controller(self._ctrl, self.running_cost, self.terminal_state_cost)
            self.set_cost(cost)
            self._cost_fp = fp
        self._refresh_dynamics_from_controller()

    def _cost_fingerprint(self):
        """Everything the cost program is built from, cheap to compare (object identities, in-place version counters,
        the few scalars): when it is unchanged since the last command the extraction, the descriptor and the device
        update (~0.2 ms of host time) are skipped."""
        c = self._ctrl
        rc, tc = self.running_cost, self.terminal_state_cost
        gc = c.goal_cost
        goal = gc.goal
        fp = [id(getattr(rc, '__func__', rc)), id(getattr(rc, '__self__', None)), id(getattr(tc, '__func__', tc)),
              id(gc.Q), getattr(gc.Q, '_version', 0), id(gc.R), getattr(gc.R, '_version', 0),
              None if goal is None else (id(goal), getattr(goal, '_version', 0)),
              float(c.terminal_cost_multiplier), id(getattr(c, 'trap_cost', None))]
        traps = c.trap_set
        if len(traps):
            tw = c.trap_set_weight
            fp.append(float(tw.reshape(-1)[0]) if (torch.is_tensor(tw) or isinstance(tw, np.ndarray)) else float(tw))
            fp.extend((id(x), getattr(x, '_version', 0), id(u), getattr(u, '_version', 0)) for x, u in traps)
        rcost = getattr(c, 'recovery_cost', None)
        if rcost is not None:
            fp.append(id(rcost))
            w = getattr(rcost, 'weights', None)
            if w is not None:
                w = w() if callable(w) else w
                fp.append(tuple(torch.as_tensor(w).reshape(-1).tolist()))
            for gs in (list(rcost.fn) if hasattr(rcost, 'fn') else [rcost]):
                g = gs.goal_set
                fp.append((id(g), getattr(g, '_version', 0), len(g), float(gs.scale)))
        return fp

    def _refresh_dynamics_from_controller(self):
        """The nominal model object can be swapped between commands (use_temp_local_nominal_model /
        use_normal_nominal_model, hybrid_model.py:189-202; online_controller.py:324-335) and an online GP changes with
        every dynamics.update (online_controller.py:85): re-read the residual whenever one is, or was, present."""
        dyn = getattr(self._ctrl, 'dynamics', None)
        model = getattr(dyn, 'nominal_model', None)
        if model is None:
            return
        online = hasattr(model, 'prior') and not hasattr(model, 'model')
        if not online:
            if self._dyn_spec.gp is not None:
                self.set_gp(None)
            return
        inner = model.prior.dyn_net
        if getattr(self, '_gp_inner', None) is not inner:
            # first time this network is seen behind a GP: it must be the one the kernels already hold
            new = extract.dynamics_from_controller(self._ctrl)
            self._gp_inner = inner
            self.set_dynamics(new)
            return
        self.set_gp(extract.gp_spec_from_model(model, self.nx, self.nu))

    # ------------------------------------------------------------------ nominal sequence
    @property
    def U(self):
        out = np.empty((self.T, self.nu))
        self._check(self._lib.tampc_mppi_get_nominal(self._h, L.dptr(out)))
        return torch.from_numpy(out)

    @U.setter
    def U(self, value):
        a = np.ascontiguousarray(torch.as_tensor(value).detach().to('cpu', torch.double).numpy().reshape(-1, self.nu))
        if a.shape[0] > self.max_horizon:
            raise ValueError("horizon {} exceeds max_horizon {}".format(a.shape[0], self.max_horizon))
        self._check(self._lib.tampc_mppi_set_nominal(self._h, L.dptr(a), a.shape[0]))
        self.T = a.shape[0]

    def change_horizon(self, horizon):
        """ExperimentalMPPI.change_horizon (controller.py:323-330): truncate, or pad with fresh noise samples."""
        U = self.U
        if horizon < self.T:
            U = U[:horizon]
        elif horizon > self.T:
            pad = self.noise_dist.sample((horizon,))
            pad[:self.T] = U
            U = pad
        self.U = U

    def reset(self):
        """MPPI.reset: U = noise_dist.sample((T,))."""
        self.U = self.noise_dist.sample((self.T,))

    # ------------------------------------------------------------------ command
    @property
    def state(self):
        """MPPI.state: the start state of the last command as a float64 tensor (materialised on access: building a
        tensor costs microseconds that the command() latency does not need to pay)."""
        return None if self._state_np is None else torch.from_numpy(self._state_np.copy())

    @state.setter
    def state(self, value):
        self._state_np = None if value is None else np.array(torch.as_tensor(value).detach().to('cpu', torch.double).numpy(),
                                                             dtype=np.float64).reshape(-1)

    def _state_array(self, state):
        if isinstance(state, np.ndarray) and state.dtype == np.float64:
            a = np.array(state, dtype=np.float64).reshape(-1)  # fast path (a private copy: `.state` must not alias the caller's buffer)
        else:
            if not torch.is_tensor(state):
                state = torch.as_tensor(np.asarray(state))
            a = np.ascontiguousarray(state.detach().to('cpu', torch.double).numpy().reshape(-1))
        if a.shape != (self.nx,):
            raise ValueError("state must have {} entries (per-sample start states are not supported)".format(self.nx))
        self._state_np = a
        return a

    def _arm_gp_draws(self, gp_eps):
        """gp_eps: (T, M, K, ny) standard normals in the reference's layout (row m*K+k of its flattened (M*K, .) batch
        at step t) -> the library's (K_local, M, T, ny)."""
        if gp_eps is None:
            return
        if self._dyn_spec.gp is None:
            raise ValueError("gp_eps given but the dynamics have no GP residual")
        e = torch.as_tensor(gp_eps).detach().to('cpu', torch.double).numpy()
        ny = self._dyn_spec.gp.resid.shape[1]
        if e.shape != (self.T, self.M, self.K, ny):
            raise ValueError("gp_eps must be (T,M,K,ny) = {}".format((self.T, self.M, self.K, ny)))
        local = np.ascontiguousarray(e[:, :, self.sample_offset:self.sample_offset + self.K_local].transpose(2, 1, 0, 3))
        self._check(self._lib.tampc_mppi_set_gp_draws(self._h, L.dptr(local), local.size))

    def command(self, state, noise=None, gp_eps=None):
        """MPPI.command(state) -> action (nu,) float64 tensor.

        noise: optional (K,T,nu) draw N(mu, Sigma) to use instead of the on-device Philox stream — the equivalent of
        patching `noise_dist.sample((K,T))` in the reference, used by the parity tests.
        gp_eps: optional (T,M,K,ny) standard normals behind the GP residual's draws (parity tests)."""
        self._refresh_from_controller()
        self._arm_gp_draws(gp_eps)
        # The state and the action go through two persistent buffers whose ctypes pointers are built once: building a
        # numpy array, its `.ctypes` view and a POINTER per command cost ~6 us of host time, as much as the whole
        # softmax update costs on the device.
        x = self._x_buf
        if isinstance(state, np.ndarray) and state.dtype == np.float64 and state.shape == x.shape:
            np.copyto(x, state)
        else:
            np.copyto(x, self._state_array(state))
        self._state_np = x
        mode, nptr, local = L.NOISE_PHILOX, None, None
        if noise is not None:
            n = torch.as_tensor(noise).detach().to('cpu', torch.double).numpy()
            if n.shape != (self.K, self.T, self.nu):
                raise ValueError("noise must be (K,T,nu) = {}".format((self.K, self.T, self.nu)))
            local = np.ascontiguousarray(n[self.sample_offset:self.sample_offset + self.K_local])
            mode, nptr = L.NOISE_INJECTED, L.dptr(local)
        if self.world_size == 1 or self._p2p:
            # (sharded over peer memory: the merge kernel does the exchange itself, nothing to do on the host)
            rc = self._lib.tampc_mppi_command(self._h, self._x_ptr, mode, nptr, self._act_ptr)
            if rc != L.OK:
                self._check(rc)
        else:
            self._command_sharded(x, mode, nptr, self._act_buf)
        out = torch.from_numpy(self._act_buf.copy())
        return out if self.d.type == 'cpu' else out.to(self.d)

    def _connect_peers(self, required=False):
        """Exchange CUDA IPC handles of the ranks' exchange buffers (torch.distributed is only the plumbing).

        The decision to use the peer exchange is COLLECTIVE: a rank whose comm_init / cudaIpcOpenMemHandle fails still
        takes part in every collective below, the ranks agree (all_reduce MIN of an ok flag) and either all of them
        use the peer exchange or all of them fall back to the NCCL all-gather (or all raise when it was required)."""
        import torch.distributed as dist
        err = None
        mine = (C.c_ubyte * L.IPC_HANDLE_BYTES)()
        try:
            self._check(self._lib.tampc_mppi_comm_init(self._h, self.rank, self.world_size, C.cast(mine, C.c_void_p)))
        except Exception as e:  # noqa: BLE001 - reported collectively below
            err = e
        handles = [None] * self.world_size
        dist.all_gather_object(handles, (err is None, bytes(mine)), group=self.group)
        if err is None and all(ok for ok, _ in handles):
            try:
                blob = (C.c_ubyte * (L.IPC_HANDLE_BYTES * self.world_size)).from_buffer_copy(b''.join(h for _, h in handles))
                self._check(self._lib.tampc_mppi_comm_connect(self._h, C.cast(blob, C.c_void_p)))
            except Exception as e:  # noqa: BLE001
                err = e
        elif err is None:
            err = RuntimeError("a peer rank could not export its exchange buffer")
        oks = [None] * self.world_size
        dist.all_gather_object(oks, err is None, group=self.group)
        self._p2p = all(oks)
        if not self._p2p:
            # every rank leaves the peer exchange together; ranks that did connect drop their mappings
            self._check(self._lib.tampc_mppi_comm_disconnect(self._h))
            if required:
                raise RuntimeError("peer exchange (exchange='p2p') could not be set up on every rank: {}".format(
                    err if err is not None else "a peer rank failed"))
        dist.barrier(group=self.group)

    def _command_sharded(self, x, mode, nptr, action):
        """K-sharded command: local rollout + partial, one all_gather of tampc_partial (NCCL over NVLink), identical
        combine on every rank."""
        if self._p2p:
            # the merge kernel does the exchange itself over peer memory: nothing to do on the host
            self._check(self._lib.tampc_mppi_command(self._h, L.dptr(x), mode, nptr, L.dptr(action)))
            return
        import torch.distributed as dist
        if self._gather is None:
            dev = torch.device('cuda', self.device_index)
            with torch.cuda.stream(self.torch_stream):
                self._gather = torch.zeros(self.world_size, L.PARTIAL_DOUBLES, dtype=torch.double, device=dev)
                self._mine = torch.zeros(L.PARTIAL_DOUBLES, dtype=torch.double, device=dev)
        self._check(self._lib.tampc_mppi_command_begin(self._h, L.dptr(x), mode, nptr, 0,
                                                       C.c_void_p(self._mine.data_ptr())))
        with torch.cuda.stream(self.torch_stream):
            dist.all_gather_into_tensor(self._gather, self._mine, group=self.group)
        self._check(self._lib.tampc_mppi_command_finish(self._h, C.c_void_p(self._gather.data_ptr()),
                                                        self.world_size, L.dptr(action)))

    # ------------------------------------------------------------------ results of the last command
    def _query(self, what, n):
        out = np.empty(n)
        self._check(self._lib.tampc_mppi_query(self._h, what, L.dptr(out), n))
        return out

    @property
    def cost_total(self):
        """Per-sample total cost of the last rollout (this rank's shard when sharded)."""
        return torch.from_numpy(self._query(L.Q_COST_TOTAL, self.K_local))

    @property
    def omega(self):
        return torch.from_numpy(self._query(L.Q_OMEGA, self.K_local))

    @property
    def perturbed_action(self):
        return torch.from_numpy(self._query(L.Q_PERTURBED, self.K_local * self.T * self.nu).reshape(self.K_local, self.T, self.nu))

    @property
    def noise(self):
        """Post-clamp noise perturbed_action - U_shifted of the last command (controller.py:394), regenerated on the
        device from the same Philox stream (or re-read from the injected draw)."""
        return torch.from_numpy(self._query(L.Q_NOISE, self.K_local * self.T * self.nu).reshape(self.K_local, self.T, self.nu))

    @property
    def actions(self):
        """ExperimentalMPPI.actions (M, K, T, nu) of the last command (controller.py:368,372): the M replicas share the
        perturbed actions, so one copy is returned with a leading axis of 1."""
        return self.perturbed_action.unsqueeze(0)

    @property
    def kernel_name(self):
        """Rollout kernel family + arithmetic a command of this handle runs (the precision argument is a request)."""
        return self._lib.tampc_mppi_kernel_name(self._h).decode()

    @property
    def stats(self):
        s = self._query(L.Q_STATS, 4)
        return {'beta': s[0], 'eta': s[1], 'argmin': int(s[2]), 'launches': int(s[3])}

    # ------------------------------------------------------------------ other reference entry points
    def get_rollouts(self, state, num_rollouts=1):
        """MPPI.get_rollouts: (num_rollouts, T, nx) states reached by the current nominal U (controller.py:434-435)."""
        x = self._state_array(state)
        out = np.empty((self.T, self.nx))
        self._check(self._lib.tampc_mppi_get_rollouts(self._h, L.dptr(x), L.dptr(out)))
        return torch.from_numpy(out).unsqueeze(0).repeat(num_rollouts, 1, 1)

    def predict_next_state(self, state, control, nominal_only=False):
        """ControllerWithModelPrediction.predict_next_state: one step from `state` under `control`, returned as a
        (1, nx) float64 ndarray like the reference.

        nominal_only=False is MPC.predict_next_state (controller.py:312-313) = `_apply_dynamics`: guard, the dynamics the
        rollouts use (incl. an installed GP residual), constrain_state.  nominal_only=True is OnlineMPC.predict_next_state
        (online_controller.py:47-58), which deliberately swaps the ORIGINAL nominal network in for the call: no GP residual,
        no guard, no constrain_state — its result feeds diff_predicted and with it the trap / local-model state machine."""
        x = self._state_array(state)
        u = np.ascontiguousarray(torch.as_tensor(control).detach().to('cpu', torch.double).numpy().reshape(-1))
        out = np.empty(self.nx)
        fn = self._lib.tampc_mppi_predict_nominal if nominal_only else self._lib.tampc_mppi_predict
        self._check(fn(self._h, L.dptr(x), L.dptr(u), L.dptr(out)))
        return out.reshape(1, -1)

    def apply_dynamics(self, state, actions):
        """OnlineMPC._apply_dynamics(state.repeat(N, 1), actions) (online_controller.py:60-79) for N candidate actions
        from ONE state, in one fused launch per K_local actions: bad-state guard, network (+ GP residual), constrain_state.
        This is the heavy step of the APF baselines' samplers (online_controller.py:583-590, 657-664).  Returns (N, nx)."""
        a = torch.as_tensor(actions).detach().to('cpu', torch.double).numpy().reshape(-1, self.nu)
        x = self._state_array(state)
        N = a.shape[0]
        out = np.empty((N, self.nx))
        cost = np.empty(self.K_local)
        states = np.empty((self.K_local, 1, self.nx))
        for lo in range(0, N, self.K_local):
            chunk = a[lo:lo + self.K_local]
            buf = np.zeros((self.K_local, 1, self.nu))
            buf[:len(chunk), 0] = chunk
            self._check(self._lib.tampc_mppi_rollout_costs(self._h, L.dptr(x), L.dptr(buf), 1, L.dptr(cost), L.dptr(states)))
            out[lo:lo + len(chunk)] = states[:len(chunk), 0]
        return torch.from_numpy(out)

    def select_action(self, state, actions):
        """The APF baselines' action selection (online_controller.py:583-590, 657-664) in one device call: candidate
        `actions` (n, nu) through `_apply_dynamics` from `state`, the installed potential on the next states, argmin on the
        device.  Returns (index, actions[index] as a float64 tensor, cost, next_state (nx,))."""
        a = np.ascontiguousarray(torch.as_tensor(actions).detach().to('cpu', torch.double).numpy().reshape(-1, self.nu))
        if a.shape[0] > self.K_local:
            raise ValueError("{} candidate actions exceed the handle's {} samples".format(a.shape[0], self.K_local))
        x = self._state_array(state)
        idx, cost = C.c_int64(), C.c_double()
        nxt = np.empty(self.nx)
        self._check(self._lib.tampc_mppi_select_action(self._h, L.dptr(x), L.dptr(a), a.shape[0], C.byref(idx),
                                                       C.cast(C.byref(cost), C.POINTER(C.c_double)), L.dptr(nxt)))
        return int(idx.value), torch.from_numpy(a[idx.value].copy()), float(cost.value), torch.from_numpy(nxt)

    def state_costs(self, states):
        """(goal_cost(x), raw trap_cost(x) with U=None) of the installed GOAL_TRAP program for (n, nx) states."""
        xs = np.ascontiguousarray(torch.as_tensor(states).detach().to('cpu', torch.double).numpy().reshape(-1, self.nx))
        g, t = np.empty(len(xs)), np.empty(len(xs))
        self._check(self._lib.tampc_mppi_state_costs(self._h, L.dptr(xs), len(xs), L.dptr(g), L.dptr(t)))
        return torch.from_numpy(g), torch.from_numpy(t)

    def normalize_trapset_cost_to_state(self, x):
        """OnlineMPPI.normalize_trapset_cost_to_state (online_controller.py:517-530): goal_cost(x) / trap_cost(x) (the trap
        cost with U=None, summed and scaled by the current trap_set_weight, controller.py:245-246), None without a trap cost
        or traps — evaluated by the device from the cost program of the bound controller."""
        self._refresh_from_controller()
        c = self._cost_spec
        if c.kind != S.COST_GOAL_TRAP or not c.use_trap_cost or len(c.trap_x) == 0:
            return None
        g, t = self.state_costs(torch.as_tensor(x).reshape(1, -1))
        return g / (t * c.trap_weight)  # MPC._trap_cost_reduce: sum over the traps * the current trap_set_weight (controller.py:245-246)

    def _compute_rollout_costs(self, perturbed_actions, state=None):
        """ExperimentalMPPI._compute_rollout_costs (controller.py:340-381): (cost_total (K,), states (M,K,T,nx),
        actions (M,K,T,nu)) for caller-given actions.  M replicas of a deterministic model are identical, so M=1
        is returned."""
        self._refresh_from_controller()
        a = np.ascontiguousarray(torch.as_tensor(perturbed_actions).detach().to('cpu', torch.double).numpy())
        K, T, nu = a.shape
        if K != self.K_local or nu != self.nu:
            raise ValueError("actions must be (K,T,nu) with K={} nu={}".format(self.K_local, self.nu))
        x = self._state_array(self.state if state is None else state)
        cost = np.empty(K)
        states = np.empty((K, T, self.nx))
        self._check(self._lib.tampc_mppi_rollout_costs(self._h, L.dptr(x), L.dptr(a), T, L.dptr(cost), L.dptr(states)))
        return torch.from_numpy(cost), torch.from_numpy(states).unsqueeze(0), torch.from_numpy(a).unsqueeze(0)

    # kernel-only timing helpers (bench.py): inputs resident in HBM, no host copies, no synchronisation
    def set_state_resident(self, state):
        x = self._state_array(state)
        self._check(self._lib.tampc_mppi_set_state_resident(self._h, L.dptr(x)))

    def command_resident(self):
        self._check(self._lib.tampc_mppi_command_resident(self._h, L.NOISE_PHILOX))

    def rollout_resident(self):
        """Only the fused rollout kernel (roofline timing)."""
        self._check(self._lib.tampc_mppi_rollout_resident(self._h, L.NOISE_PHILOX))

    def synchronize(self):
        self._check(self._lib.tampc_mppi_synchronize(self._h))


def b200_controller_class(base, **b200_options):
    """Subclass of a reference controller class derived from MPPI_MPC (OnlineMPPI, MPPI_MPC itself) whose own line
    `self.mpc = ExperimentalMPPI(self._apply_dynamics, self._running_cost, self.nx, ...)` (controller.py:418-421) builds a
    B200MPPI: nothing else of the reference is overridden or edited.  `b200_options` go to B200MPPI (precision, cuda_device,
    seed, rank / world_size / process_group, ...)."""
    import functools
    import sys
    owner = next((c for c in base.__mro__ if 'ExperimentalMPPI' in vars(sys.modules[c.__module__]) and c.__name__ == 'MPPI_MPC'), None)
    if owner is None:
        raise TypeError("{} does not derive from tampc's MPPI_MPC".format(base.__name__))
    mod = sys.modules[owner.__module__]

    class B200Controller(base):
        def __init__(self, *args, **kwargs):
            original = mod.ExperimentalMPPI
            mod.ExperimentalMPPI = functools.partial(B200MPPI, **b200_options)
            try:
                super().__init__(*args, **kwargs)
            finally:
                mod.ExperimentalMPPI = original
            _hook_controller(self, self.mpc)

    B200Controller.__name__ = 'B200' + base.__name__
    B200Controller.__qualname__ = B200Controller.__name__
    return B200Controller


def _original_nominal_network(ctrl):
    """The network OnlineMPC.predict_next_state evaluates (online_controller.py:47-58): dynamics._original_nominal_model,
    or its prior.dyn_net if that is an online model.  None for controllers without a hybrid dynamics object."""
    dyn = getattr(ctrl, 'dynamics', None)
    orig = getattr(dyn, '_original_nominal_model', None)
    if orig is None:
        return None
    if hasattr(orig, 'prior') and not hasattr(orig, 'model'):
        orig = getattr(orig.prior, 'dyn_net', None)
    return orig


def install(ctrl, predict_on_device=True, **kwargs):
    """Swap `ctrl.mpc` (an ExperimentalMPPI built at controller.py:418-421) for a B200MPPI bound to `ctrl`.

    The controller's own state machine (trap detection, recovery, MAB; online_controller.py:437-515) is untouched:
    it keeps assigning `ctrl.mpc.running_cost / terminal_state_cost`, calling `change_horizon` and `command`.

    predict_on_device: also serve `ctrl.predict_next_state` (called once per control step, controller.py:306) with one
    30 us launch.  For an OnlineMPC that is the ORIGINAL NOMINAL network only (no GP residual, guard or constrain_state,
    online_controller.py:47-58) and only if that network is the one the kernels hold; for a plain MPC it is
    `_apply_dynamics` (controller.py:312-313).  Otherwise the controller's own host method stays in place."""
    mpc = B200MPPI.from_controller(ctrl, **kwargs)
    ctrl.mpc = mpc
    if predict_on_device:
        _hook_controller(ctrl, mpc)
    return mpc


def _hook_controller(ctrl, mpc):
    """Serve the controller's per-step model queries from the device as well (see install)."""
    if hasattr(ctrl, 'predict_next_state'):
        dyn = getattr(ctrl, 'dynamics', None)
        if hasattr(dyn, '_original_nominal_model'):
            cur = getattr(dyn, 'nominal_model', None)
            held = cur.prior.dyn_net if (hasattr(cur, 'prior') and not hasattr(cur, 'model')) else cur
            if _original_nominal_network(ctrl) is held:
                ctrl.predict_next_state = lambda state, control: mpc.predict_next_state(state, control, nominal_only=True)
        else:
            ctrl.predict_next_state = mpc.predict_next_state
    if hasattr(ctrl, 'normalize_trapset_cost_to_state'):
        ctrl.normalize_trapset_cost_to_state = mpc.normalize_trapset_cost_to_state  # online_controller.py:471,517-530
