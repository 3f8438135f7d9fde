"""TEST INFRASTRUCTURE: a duck-typed stand-in for a live tampc OnlineMPPI controller, rebuilt from a golden
RolloutSpec, so that the integration path (extract.spec_from_controller -> B200MPPI.from_controller -> install ->
per-command cost refresh, recovery swap) runs on the GPU box, where /root/reference does not exist.

It mirrors only the attribute surface extract.py reads: the env callables are real functions with the
reference's semantics (block_push.py:954-986, 663-666; push_main.py:190-193), the networks are real torch modules,
the scalers are the oracle shim's classes.  Nothing here is used by the product."""
import math
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim'))
from arm_pytorch_utilities import preprocess  # noqa: E402  (shim restatement)

from tampc_b200 import spec as S  # noqa: E402


def _t(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.double)


def _sequential(m: S.MLPSpec):
    layers = []
    n = len(m.weights)
    for i, (w, b) in enumerate(zip(m.weights, m.biases)):
        lin = torch.nn.Linear(w.shape[1], w.shape[0]).double()
        with torch.no_grad():
            lin.weight.copy_(_t(w))
            lin.bias.copy_(_t(b))
        if i + 1 < n:
            layers += [lin, torch.nn.LeakyReLU(m.slope)]
        else:
            layers.append(torch.nn.Sequential(lin))
    return torch.nn.Sequential(*layers)


def _scaler(a):
    if a is None:
        return preprocess.NullSingleTransformer()
    s = preprocess.MinMaxScaler()
    s._scale, s._min = _t(a.scale), _t(a.offset)
    return s


class AlwaysSelectNominal:
    def sample_class(self, state, action, *args):
        return torch.zeros((state.shape[0],), dtype=torch.long)


class _GoalSetCost:
    def __init__(self, goal_set, scale):
        self.goal_set, self.scale, self.goal_weights = goal_set, scale, None
        self.reduce = min_cost


def min_cost(costs):
    return costs.min(dim=0).values


class _ComposeCost:
    def __init__(self, fn, weights):
        self.fn, self.weights = fn, weights


def make_controller(spec: S.RolloutSpec, U):
    d, c, s = spec.dynamics, spec.cost, spec.sampler
    assert d.kind in (S.DYN_REX, S.DYN_MLP)
    ctrl = types.SimpleNamespace()
    ctrl.nx, ctrl.nu = d.nx, d.nu

    def compare_to_goal(state, goal):
        if goal.dim() == 1:
            goal = goal.view(1, -1)
        diff = state - goal
        for k in c.ang_dims:
            col = diff[:, k].clone()
            col[col > math.pi] -= 2 * math.pi
            col[col < -math.pi] += 2 * math.pi
            diff[:, k] = col
        return diff

    def state_dist(diff):
        return (diff * _t(c.dist_scale)).norm(dim=1)

    def u_similarity(u1, u2):
        if u2.dim() == 1:
            u2 = u2.view(1, -1)
        dims = list(c.usim_dims)
        return torch.cosine_similarity(u1[:, dims], u2[:, dims], dim=-1).clamp(0, 1)

    def constrain_state(state):
        for k in d.wrap_dims:
            state[:, k] = torch.remainder(state[:, k] + math.pi, 2 * math.pi) - math.pi
        return state

    ctrl.compare_to_goal, ctrl.state_dist, ctrl.constrain_state = compare_to_goal, state_dist, constrain_state
    ctrl.gating = AlwaysSelectNominal()
    # ---- dynamics: NetworkModelWrapper-like object behind a HybridDynamicsModel-like object
    net = _sequential(d.net)
    if d.kind == S.DYN_REX:
        inv = types.SimpleNamespace(encoder=types.SimpleNamespace(model=_sequential(d.encoder)),
                                    extracted_linear_decoder=types.SimpleNamespace(model=_sequential(d.decoder)),
                                    x_extractor=torch.nn.Linear(d.nx, d.ext_w.shape[0]).double(), nv=d.nv,
                                    xu_to_z=lambda *a: None, get_dx=lambda *a: None)
        with torch.no_grad():
            inv.x_extractor.weight.copy_(_t(d.ext_w))
            inv.x_extractor.bias.copy_(_t(d.ext_b))
        tsf = preprocess.Compose([
            preprocess.PytorchTransformer(preprocess.NullSingleTransformer(), _scaler(d.aff_y)),
            types.SimpleNamespace(tsf=inv),
            preprocess.PytorchTransformer(_scaler(d.aff_in), _scaler(d.aff_out))])
    else:
        tsf = preprocess.PytorchTransformer(_scaler(d.aff_in), _scaler(d.aff_out))
    cfg = types.SimpleNamespace(predict_difference=True, predict_all_dims=True)
    ds = types.SimpleNamespace(preprocessor=types.SimpleNamespace(tsf=tsf), original_config=lambda: cfg)
    nominal = types.SimpleNamespace(model=net, ds=ds)
    ctrl.dynamics = types.SimpleNamespace(nominal_model=nominal, _original_nominal_model=nominal, num_local_models=lambda: 0)  # hybrid_model.py:113
    # ---- cost wiring of MPC.__init__ (controller.py:216-230)
    ctrl.goal_cost = types.SimpleNamespace(Q=_t(c.Q) if c.kind == S.COST_GOAL_TRAP else torch.eye(d.nx, dtype=torch.double),
                                           R=_t(c.R) if c.kind == S.COST_GOAL_TRAP else torch.eye(d.nu, dtype=torch.double),
                                           goal=_t(c.goal) if c.kind == S.COST_GOAL_TRAP else torch.zeros(d.nx, dtype=torch.double))
    ctrl.trap_set = []
    ctrl.trap_set_weight = 1.0
    ctrl.terminal_cost_multiplier = 50.0
    ctrl.trap_cost = types.SimpleNamespace(u_similarity=u_similarity, goal_weights=None)
    ctrl.recovery_cost = None
    if c.kind == S.COST_GOAL_TRAP:
        ctrl.trap_set = [(_t(x), _t(u)) for x, u in zip(c.trap_x, c.trap_u)]
        ctrl.trap_set_weight = c.trap_weight
        if not c.use_trap_cost:
            ctrl.trap_cost = None
        if c.terminal_multiplier is not None:
            ctrl.terminal_cost_multiplier = c.terminal_multiplier

    class _Methods:  # bound methods, so that identity checks work like on the reference controller
        def _running_cost(self, state, action):
            raise RuntimeError("host cost callback must never run on the fused path")

        def _terminal_cost(self, state, action):
            raise RuntimeError("host cost callback must never run on the fused path")

        def _recovery_running_cost(self, state, action):
            raise RuntimeError("host cost callback must never run on the fused path")

    meth = _Methods()
    ctrl._running_cost, ctrl._terminal_cost = meth._running_cost, meth._terminal_cost
    ctrl._recovery_running_cost = meth._recovery_running_cost
    # ---- the ExperimentalMPPI object MPPI_MPC builds (controller.py:418-421)
    mpc = types.SimpleNamespace(K=s.K, T=s.T, nu=s.nu, lambda_=s.lambda_, noise_mu=_t(s.noise_mu),
                                noise_sigma=_t(s.noise_sigma), u_min=None if s.u_min is None else _t(s.u_min),
                                u_max=None if s.u_max is None else _t(s.u_max), u_init=_t(s.u_init),
                                sample_null_action=s.sample_null_action, M=s.rollout_samples,
                                rollout_var_cost=s.rollout_var_cost, rollout_var_discount=s.rollout_var_discount,
                                U=_t(U).clone(), running_cost=ctrl._running_cost,
                                terminal_state_cost=ctrl._terminal_cost if (c.kind == S.COST_GOAL_TRAP and c.terminal_multiplier is not None) else None)
    if c.kind == S.COST_RECOVERY:
        sets = [_GoalSetCost(_t(g), c.recovery_scale) for g in c.goal_sets]
        w = _t(c.set_weights)
        ctrl.recovery_cost = _ComposeCost(sets, lambda: w) if len(sets) > 1 else sets[0]
        mpc.running_cost = ctrl._recovery_running_cost
        mpc.terminal_state_cost = None
    ctrl.mpc = mpc
    ctrl.predict_next_state = lambda state, control: None
    return ctrl


class RBFKernel:
    def __init__(self, lengthscale):
        self.lengthscale = _t(lengthscale).view(-1, 1, 1)


def make_online_gp(spec: S.RolloutSpec, nominal):
    """Duck-typed OnlineGPMixing (online_model.py:324-457) around `nominal`, frozen at spec.dynamics.gp: what
    hybrid_model.use_temp_local_nominal_model (GP_ORIGINAL) or nominal_adapt=GP_KERNEL_INDEP_OUT (GP_AT_NETWORK)
    leave in `ctrl.dynamics.nominal_model`."""
    g = spec.dynamics.gp
    n, ny = g.resid.shape
    if g.space == S.GP_ORIGINAL:
        pre = types.SimpleNamespace(tsf=preprocess.PytorchTransformer(_scaler(g.aff_in), _scaler(g.aff_out)))
    else:
        pre = nominal.ds.preprocessor
    gp = types.SimpleNamespace(
        training=False,
        covar_module=types.SimpleNamespace(outputscale=_t(g.outputscale), base_kernel=RBFKernel(g.lengthscale)),
        # mean function at the training inputs; this stand-in keeps y - m(X) in `y` and answers m(X) = 0
        forward=lambda xu: types.SimpleNamespace(mean=torch.zeros(len(xu), ny, dtype=torch.double)))
    lik = types.SimpleNamespace(task_noises=_t(g.noise) * 0.75, noise=_t(g.noise[:1]) * 0 + _t(g.noise) * 0.25)
    return types.SimpleNamespace(prior=types.SimpleNamespace(dyn_net=nominal), ds=types.SimpleNamespace(preprocessor=pre),
                                 gp=gp, likelihood=lik, xu=_t(g.X), y=_t(g.resid), use_independent_outputs=True,
                                 sample_dynamics=g.sample)


def identity_preprocessor():
    """A preprocessor whose x side cannot be inverted (like the invariant transform, invariant.py:642-643)."""
    def invert_x(x):
        raise RuntimeError("Transform loses information and cannot invert X")
    return types.SimpleNamespace(invert_x=invert_x, transform_y=lambda y: y, tsf=None)


def fake_online_model_module():
    """Stand-in for the module tampc.dynamics.online_model on the GPU box (no reference tree there): the abstract base, the
    refit-strategy enum and the DATA BOOKKEEPING of OnlineGPMixing that B200OnlineGP inherits — constructor attributes,
    reset, the roll / append window and the refit-strategy dispatch of `_update` (online_model.py:317-410) — behaviourally
    restated for the test; where the reference exists, tests/test_reference_live.py runs the same subclass over the real
    class."""
    import enum

    class RefitGPStrategy(enum.IntEnum):
        RESET_DATA = 0
        RECREATE_GP = 1
        RECREATE_ALL = 2

    class OnlineDynamicsModel:
        def __init__(self, ds, state_difference, device=torch.device('cpu')):
            self.ds, self.state_difference, self.d, self.dtype = ds, state_difference, device, torch.double
            self.nx, self.nu = ds.config.nx, ds.config.nu

    class OnlineGPMixing(OnlineDynamicsModel):
        def __init__(self, prior, ds, state_difference, max_data_points=50, training_iter=100, slice_to_use=None,
                     refit_strategy=RefitGPStrategy.RESET_DATA, use_independent_outputs=False, allow_update=True, sample=True, **kw):
            super().__init__(ds, state_difference, **kw)
            self.prior, self.max_data_points, self.training_iter = prior, max_data_points, training_iter
            self.refit_strategy, self.use_independent_outputs = refit_strategy, use_independent_outputs
            self.allow_update, self.sample_dynamics = allow_update, sample
            self.xu = self.y = self.likelihood = self.gp = self.optimizer = self.last_prediction = None
            self.last_loss = self.last_loss_diff = 0
            xu, y, _ = ds.training_set()
            sl = slice(max_data_points) if slice_to_use is None else slice_to_use
            self.init_xu, self.init_y = xu[sl], y[sl]
            self.reset()

        def reset(self):
            self.xu, self.y = self.init_xu.clone(), self.init_y.clone()
            self._recreate_all()
            if len(self.xu):
                self._fit_params(self.training_iter)

        def _update(self, px, pu, y):
            self.last_prediction = None
            if not self.allow_update:
                return
            row, y = torch.cat((px.view(1, -1), pu.view(1, -1)), dim=1), y.view(1, -1)
            if self.xu.shape[0] < self.max_data_points:
                self.xu, self.y = torch.cat((self.xu, row)), torch.cat((self.y, y))
            else:
                self.xu, self.y = torch.roll(self.xu, -1, dims=0), torch.roll(self.y, -1, dims=0)
                self.xu[-1], self.y[-1] = row, y
            if self.refit_strategy is RefitGPStrategy.RESET_DATA:
                self.gp.set_train_data(self.xu, self.y, strict=False)
                self._fit_params(self.training_iter // 10)
            elif self.refit_strategy is RefitGPStrategy.RECREATE_GP:
                self._recreate_gp()
                self._fit_params(self.training_iter // 3)
            else:
                self._recreate_all()
                self._fit_params(self.training_iter)

    return types.SimpleNamespace(OnlineDynamicsModel=OnlineDynamicsModel, OnlineGPMixing=OnlineGPMixing, RefitGPStrategy=RefitGPStrategy)
