"""CPU, only where /root/reference exists: spec extraction from the reference's live controllers and the oracle
against the reference's own classes on fresh inputs (the committed goldens cover the GPU box)."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import ref_harness as rh  # noqa: E402

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference tree not present (GPU box)")

from oracle import mppi_oracle as O  # noqa: E402
from tampc_b200 import extract, spec as S  # noqa: E402


@pytest.fixture(scope='module')
def push_ctrl():
    torch.manual_seed(0)
    ctrl, ds, env = rh.build_task('push', num_samples=96, rollout_samples=2, horizon=15)
    ctrl.set_goal(np.array([0.55, -0.65, 0., 0., 0.]))
    return ctrl


def _run_ref(ctrl, spec, state, U0, noise):
    mpc = ctrl.mpc
    mpc.U = U0.clone()
    orig = mpc.noise_dist.sample
    mpc.noise_dist.sample = lambda shape=(): noise.clone() if tuple(shape) == (spec.sampler.K, spec.sampler.T) else orig(shape)
    try:
        with torch.no_grad():
            return mpc.command(state.clone())
    finally:
        mpc.noise_dist.sample = orig


def test_extraction_reads_the_reference_objects(push_ctrl):
    spec = extract.spec_from_controller(push_ctrl)
    d, c = spec.dynamics, spec.cost
    assert d.kind == S.DYN_REX and d.macs_per_step == 3523 and d.wrap_dims == (2,) and d.bad_state_norm == 1e5
    assert d.encoder.dims == [8, 16, 32, 5] and d.net.dims == [5, 32, 32, 5] and d.decoder.dims == [2, 16, 32, 25]
    assert c.ang_dims == (2,) and c.usim_dims == (0, 2) and c.terminal_multiplier == 50.0
    np.testing.assert_allclose(c.dist_scale, [1, 1, math.sqrt(0.3 ** 2 * 2 / 12), 0, 0])
    assert spec.sampler.lambda_ == 0.01 and spec.sampler.rollout_samples == 2


@pytest.mark.parametrize('seed', [3, 4])
def test_oracle_matches_live_reference_on_fresh_inputs(push_ctrl, seed):
    ctrl = push_ctrl
    g = torch.Generator().manual_seed(seed)
    ctrl.trap_set.clear()
    for _ in range(seed):
        ctrl.trap_set.append((torch.randn(5, dtype=torch.double, generator=g) * 0.5,
                              torch.randn(3, dtype=torch.double, generator=g)))
    ctrl.trap_set_weight = 0.1 * seed
    spec = extract.spec_from_controller(ctrl)
    state = torch.randn(5, dtype=torch.double, generator=g) * 0.4
    U0 = torch.rand(spec.sampler.T, 3, dtype=torch.double, generator=g) * 0.5
    noise = O.sample_noise(spec, seed)
    a_ref = _run_ref(ctrl, spec, state, U0, noise)
    out = O.command(spec, state, U0, noise)
    rel = ((out['cost_total'] - ctrl.mpc.cost_total).abs() / ctrl.mpc.cost_total.abs()).max().item()
    assert rel < 1e-12 and out['argmin'] == int(ctrl.mpc.cost_total.argmin())
    assert (out['action'] - a_ref).abs().max().item() < 1e-12


def test_unrecognised_callbacks_are_errors_not_fallbacks(push_ctrl):
    ctrl = push_ctrl
    with pytest.raises(extract.UnsupportedSpec):
        extract.cost_from_controller(ctrl, running_cost=lambda x, u: x.sum(dim=1))
    with pytest.raises(extract.UnsupportedSpec):
        extract.probe_dist_scale(lambda d: d.abs().sum(dim=1), 5)
    with pytest.raises(extract.UnsupportedSpec):
        extract.probe_wrap_dims(lambda s: s * 2, 5)
    with pytest.raises(extract.UnsupportedSpec):
        extract.mlp_from_module(torch.nn.Sequential(torch.nn.Linear(3, 3), torch.nn.Tanh(), torch.nn.Linear(3, 1)))


def test_peg_env_functions_probe_correctly():
    ctrl, ds, env = rh.build_task('peg', num_samples=16, rollout_samples=2)
    ctrl.set_goal(np.array([0, 0.2, 0.19, 0, 0]))
    spec = extract.spec_from_controller(ctrl)
    assert spec.cost.ang_dims == () and spec.cost.usim_dims == (0, 1) and spec.dynamics.wrap_dims == ()
    np.testing.assert_array_equal(spec.cost.dist_scale, [1, 1, 0, 0, 0])
    assert spec.dynamics.macs_per_step == 3507


def test_oracle_follows_the_live_online_gp_through_updates():
    """The TAMPC flow (online_controller.py:324-335, 85): the nominal model becomes a temporary local GP fed one
    transition per control step.  After each update the re-extracted spec + oracle must reproduce the reference's
    command with the action noise and the GP draws injected."""
    torch.manual_seed(1)
    ctrl, ds, env = rh.build_task('push', num_samples=24, rollout_samples=4, horizon=8)
    ctrl.set_goal(np.array([0.55, -0.65, 0., 0., 0.]))
    g = torch.Generator().manual_seed(11)
    x = torch.tensor([0.6, 0.14, -1.2, 5., 12.], dtype=torch.double)
    with torch.enable_grad():
        ctrl.dynamics.use_temp_local_nominal_model()
    for step in range(3):
        u = torch.rand(3, dtype=torch.double, generator=g) * 2 - 1
        nx_ = x + 0.004 * torch.randn(5, dtype=torch.double, generator=g) * torch.tensor([1, 1, 1, 40, 40.])
        with torch.enable_grad():
            ctrl.dynamics.update(x, u, nx_)
        x = nx_
        spec = extract.spec_from_controller(ctrl)
        assert spec.dynamics.gp.space == S.GP_ORIGINAL and len(spec.dynamics.gp.X) == step + 1
        U0 = torch.rand(spec.sampler.T, 3, dtype=torch.double, generator=g) * 0.5
        noise = O.sample_noise(spec, 20 + step)
        eps = torch.randn(spec.sampler.T, 4, spec.sampler.K, 5, dtype=torch.double, generator=g)
        with rh.inject_gp_draws([eps[t].reshape(-1, 5) for t in range(spec.sampler.T)]) as inj:
            a_ref = _run_ref(ctrl, spec, x, U0, noise)
            assert not inj.draws
        out = O.command(spec, x, U0, noise, eps)
        rel = ((out['cost_total'] - ctrl.mpc.cost_total).abs() / ctrl.mpc.cost_total.abs()).max().item()
        assert rel < 1e-10 and out['argmin'] == int(ctrl.mpc.cost_total.argmin())
        assert (out['action'] - a_ref).abs().max().item() < 1e-11
        # OnlineMPC.predict_next_state (online_controller.py:47-58) ignores the installed GP: it evaluates the ORIGINAL
        # nominal network, without guard or constrain_state — what tampc_mppi_predict_nominal serves on the device
        nominal_only = O.predict(_without_gp(spec.dynamics), x.view(1, -1), u.view(1, -1))
        with torch.no_grad():
            got = torch.from_numpy(ctrl.predict_next_state(x.clone(), u.clone()))
        assert (got - nominal_only).abs().max().item() < 1e-12
        with_gp, _ = O.apply_dynamics(spec.dynamics, x.view(1, -1), u.view(1, -1), torch.zeros(1, 5, dtype=torch.double))
        assert (with_gp - nominal_only).abs().max().item() > 1e-9
        from tampc_b200.mppi import _original_nominal_network
        assert _original_nominal_network(ctrl) is ctrl.dynamics.nominal_model.prior.dyn_net
    # leaving non-nominal dynamics restores the plain network (hybrid_model.py:201-202)
    ctrl.dynamics.use_normal_nominal_model()
    assert extract.spec_from_controller(ctrl).dynamics.gp is None


def _without_gp(d):
    import copy
    d = copy.copy(d)
    d.gp = None
    return d


def test_device_fitted_gp_class_is_recognised_by_the_reference():
    """install_gp swaps OnlineGPMixing for a subclass of the reference's own OnlineDynamicsModel: `update()` stays the
    reference's code and HybridDynamicsModel._uses_local_model_api (hybrid_model.py:185-187) accepts instances."""
    ref = rh.load_reference()
    from tampc_b200.online_gp import make_online_gp_class
    cls = make_online_gp_class(ref.online_model)
    assert issubclass(cls, ref.online_model.OnlineGPMixing) and issubclass(cls, ref.online_model.OnlineDynamicsModel)
    assert cls.update is ref.online_model.OnlineDynamicsModel.update
    # the data bookkeeping is the reference's own code, only the gpytorch-facing pieces are replaced
    for name in ('reset', '_update'):
        assert getattr(cls, name) is getattr(ref.online_model.OnlineGPMixing, name)
    for name in ('_recreate_all', '_recreate_gp', '_fit_params', '_make_prediction'):
        assert getattr(cls, name) is not getattr(ref.online_model.OnlineGPMixing, name)
    assert not getattr(cls, '__abstractmethods__', None)


def test_apf_potentials_and_trapset_normalisation_match_the_reference_classes(push_ctrl):
    """the oracle's restatement of the APF baselines' potentials (ArtificialRepulsionCost, APFHelicoid2D,
    goal cost with U=None; cost.py:11-18,236-289 as used at online_controller.py:561-590,639-664) and of
    OnlineMPPI.normalize_trapset_cost_to_state (online_controller.py:517-530) against the reference's own classes, through
    the extraction the device path uses (extract.apf_cost_from_controller)."""
    import types
    ref = rh.load_reference()
    ctrl = push_ctrl
    g = torch.Generator().manual_seed(5)
    X = torch.randn(257, 5, dtype=torch.double, generator=g) * torch.tensor([0.6, 0.6, 2.5, 3., 3.])
    traps = [torch.randn(5, dtype=torch.double, generator=g) * 0.4 for _ in range(4)]
    traps.append(X[7].clone() + 1e-9)  # a candidate sitting on a repulsive state: clamped at TRAP_MAX_COST
    rep = ref.cost.ArtificialRepulsionCost(traps, ctrl.compare_to_goal, ctrl.state_dist, 0.7, gain=2.5)
    apfvo = types.SimpleNamespace(nx=5, nu=3, compare_to_goal=ctrl.compare_to_goal, state_dist=ctrl.state_dist, goal_cost=ctrl.goal_cost,
                                  cost=ref.cost.ComposeCost([ctrl.goal_cost, rep]))
    c = extract.apf_cost_from_controller(apfvo)
    assert c.kind == S.COST_APF and c.apf_use_goal and len(c.trap_x) == 5 and c.apf_max_dist == 0.7 and c.apf_gain == 2.5
    want = apfvo.cost(X.clone())
    got = O.apf_cost(c, X.clone())
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12) and int(got.argmin()) == int(want.argmin())
    # APFSP: the helicoid around the active obstacle, both senses, and the plain goal potential
    for cw in (False, True):
        hel = ref.cost.APFHelicoid2D(torch.tensor([0.1, -0.2], dtype=torch.double), torch.tensor([0.4, 0.3], dtype=torch.double), clockwise=cw)
        ch = extract.apf_cost_from_controller(apfvo, hel)
        assert not ch.apf_use_goal and len(ch.trap_x) == 0
        assert torch.allclose(O.apf_cost(ch, X.clone()), hel(X.clone()), rtol=1e-12, atol=1e-12)
    cg = extract.apf_cost_from_controller(apfvo, ctrl.goal_cost)
    assert torch.allclose(O.apf_cost(cg, X.clone()), ctrl.goal_cost(X.clone()), rtol=1e-12, atol=1e-12)
    with pytest.raises(extract.UnsupportedSpec):
        extract.apf_cost_from_controller(apfvo, ref.cost.CostLeaveState(X[0], 1., 1., ctrl.compare_to_goal, 100.))
    # normalize_trapset_cost_to_state on the live OnlineMPPI
    ctrl.trap_set.clear()
    assert ctrl.normalize_trapset_cost_to_state(X[0]) is None
    for k in range(3):
        ctrl.trap_set.append((torch.randn(5, dtype=torch.double, generator=g) * 0.5, torch.randn(3, dtype=torch.double, generator=g)))
    spec = extract.spec_from_controller(ctrl)
    for i in range(5):
        want = ctrl.normalize_trapset_cost_to_state(X[i].clone())
        got, _, _ = O.normalize_trapset_cost_to_state(spec.cost, X[i])
        assert abs(float(got) - float(want)) <= 1e-12 * abs(float(want))


def test_real_reference_controller_drives_the_drop_in_through_its_own_constructor_line(monkeypatch):
    """The join: the reference's REAL OnlineMPPI — unmodified, its own `self.mpc = ExperimentalMPPI(
    self._apply_dynamics, self._running_cost, self.nx, ...)` line (controller.py:418-421) — builds a B200MPPI through the
    pytorch_mppi keyword constructor (b200_controller_class), and its own command() loop (controller.py:286-307,
    online_controller.py:81-90,437-515) drives it: cost program read at the first command (set_goal comes after
    construction), re-read when the trap set / weights / recovery cost change, horizon changes, nominal-only
    predict_next_state.  On this CPU box the C ABI is answered by the oracle (tests/fake_lib.py); everything up to the ABI
    call is the product's host code.  The same episode on the unmodified reference controller must give the same actions."""
    import fake_lib
    from tampc_b200 import _lib as L
    from tampc_b200.mppi import B200MPPI, b200_controller_class
    ref = rh.load_reference()
    fake = fake_lib.FakeLib()
    monkeypatch.setattr(L, 'load', lambda: fake)
    cls = b200_controller_class(ref.online_controller.OnlineMPPI, precision='tf32x3')
    assert issubclass(cls, ref.online_controller.OnlineMPPI) and cls._compute_action is ref.online_controller.OnlineMPPI._compute_action
    torch.manual_seed(3)
    cb, _, _ = rh.build_task('push', num_samples=64, rollout_samples=2, horizon=12, controller_class=cls)
    torch.manual_seed(3)
    cr, _, _ = rh.build_task('push', num_samples=64, rollout_samples=2, horizon=12)
    assert isinstance(cb.mpc, B200MPPI) and not isinstance(cr.mpc, B200MPPI)
    fake.bind(cb.mpc)
    assert fake.calls[:2] == ['create', 'set_model'] and fake.model_desc.kind == S.DYN_REX and fake.cfg.num_samples == 64
    # the attribute surface the controllers touch
    assert cb.mpc.T == cr.mpc.T == 12 and cb.original_horizon == 12 and cb.mpc.K == 64 and cb.mpc.M == 2
    assert torch.equal(cb.mpc.U, cr.mpc.U)  # same torch RNG stream as pytorch_mppi's U = noise_dist.sample((T,))
    assert torch.equal(cb.mpc.noise_sigma, cr.mpc.noise_sigma) and cb.mpc.lambda_ == cr.mpc.lambda_
    goal = np.array([0.55, -0.65, 0., 0., 0.])
    cb.set_goal(goal)
    cr.set_goal(goal)
    g = torch.Generator().manual_seed(17)
    obs = np.array([-0.4, 0.23, -0.6, 0., 0.])
    spec_s = cb.mpc.spec.sampler

    def step(obs, T):
        noise = O.sample_noise(cb.mpc.spec, int(torch.randint(1 << 30, (1,), generator=g)), K=64, T=T)
        fake.next_noise = noise.clone()
        orig = cr.mpc.noise_dist.sample
        cr.mpc.noise_dist.sample = lambda shape=(): noise.clone() if tuple(shape) == (64, T) else orig(shape)
        try:
            with torch.no_grad():
                ur = cr.command(obs.copy())
        finally:
            cr.mpc.noise_dist.sample = orig
        ub = cb.command(obs.copy())
        assert 'command' in fake.calls
        np.testing.assert_allclose(ub, ur, rtol=0, atol=1e-12)
        np.testing.assert_allclose(cb.predicted_next_state, cr.predicted_next_state, rtol=0, atol=1e-12)
        np.testing.assert_allclose(cb.mpc.U.numpy(), cr.mpc.U.numpy(), rtol=0, atol=1e-12)
        # the world follows the nominal model: the controllers stay in nominal dynamics (no GP fit in this test)
        obs[:] = cr.predicted_next_state.reshape(-1)
        return ub

    for i in range(4):
        u = step(obs, 12)
    assert fake.calls.count('set_cost') == 2  # the placeholder at construction, then the real program at the first command
    assert fake.cost_desc.kind == S.COST_GOAL_TRAP and fake.cost_desc.has_terminal == 1 and abs(fake.cost_desc.goal[1] + 0.65) < 1e-15
    assert 'predict_nominal' in fake.calls and 'predict' not in fake.calls  # OnlineMPC.predict_next_state: nominal network only
    # a trap appears and its weight is tuned (online_controller.py:353,461-478): the next command must carry it
    for c in (cb, cr):
        c.trap_set.append((torch.tensor([-0.3, 0.2, -0.5, 0., 0.], dtype=torch.double), torch.tensor([0.2, 0.8, -0.1], dtype=torch.double)))
        c.trap_set_weight = 0.37
    n_before = fake.calls.count('set_cost')
    u = step(obs, 12)
    # (the controller's own annealing has already acted on the weight: 0.37 * trap_cost_annealing_rate, online_controller.py:469-478)
    assert fake.calls.count('set_cost') == n_before + 1 and fake.cost_desc.n_traps == 1
    assert fake.cost_desc.trap_weight == float(cr.trap_set_weight) == float(cb.trap_set_weight) != 1.0
    w = cb.normalize_trapset_cost_to_state(torch.from_numpy(obs))
    assert abs(float(w) - float(cr.normalize_trapset_cost_to_state(torch.from_numpy(obs)))) < 1e-12 and 'state_costs' in fake.calls
    # recovery mode exactly as _start_recovery_mode / _end_recovery_mode leave the mpc (online_controller.py:370-402,416-418)
    for c in (cb, cr):
        gs = [torch.tensor([-0.45, 0.2, -0.6, 0., 0.], dtype=torch.double), torch.tensor([-0.5, 0.25, -0.55, 0., 0.], dtype=torch.double)]
        c.recovery_cost = ref.cost.ComposeCost(
            [ref.cost.GoalSetCost(gs, c.compare_to_goal, c.state_dist, reduce=ref.cost.min_cost, scale=c.recovery_scale),
             ref.cost.GoalSetCost(gs[:1], c.compare_to_goal, c.state_dist, reduce=ref.cost.min_cost, scale=c.recovery_scale)],
            weights=torch.tensor([0.7, 0.3], dtype=torch.double))
        c.mpc.running_cost = c._recovery_running_cost
        c.mpc.terminal_state_cost = None
    torch.manual_seed(5)
    cb.mpc.change_horizon(5)
    torch.manual_seed(5)
    cr.mpc.change_horizon(5)
    u = step(obs, 5)
    assert fake.cost_desc.kind == S.COST_RECOVERY and fake.cost_desc.n_sets == 2 and fake.T == 5
    for c in (cb, cr):
        c.mpc.running_cost = c._running_cost
        c.mpc.terminal_state_cost = c._terminal_cost
    torch.manual_seed(6)
    cb.mpc.change_horizon(12)   # pads U with fresh noise_dist samples (controller.py:323-330): same RNG stream on both
    torch.manual_seed(6)
    cr.mpc.change_horizon(12)
    np.testing.assert_array_equal(cb.mpc.U.numpy(), cr.mpc.U.numpy())
    step(obs, 12)
    assert fake.cost_desc.kind == S.COST_GOAL_TRAP
    with torch.no_grad():
        np.testing.assert_allclose(cb.get_rollouts(obs), cr.get_rollouts(obs), rtol=0, atol=1e-12)


@pytest.mark.parametrize('enc_h,dyn_h,dec_h', [((24, 40), (48, 24), (20, 36)), ((64,), (32, 32, 16), (16,))])
def test_networks_of_other_shapes_extract_and_match_the_live_reference(enc_h, dyn_h, dec_h):
    """The reference builds its networks from `encoder_opts` / `decoder_opts` / `dynamics_opts` h_units
    (invariant.py:745-785, util.py:560): swap the live controller's three networks for ones of other widths and depths made
    by the same `make_sequential_network`, and the extraction + oracle still reproduce the reference's own command.  (On the
    device such shapes run on the run-time-shaped kernel, held to this oracle by tests/test_gpu_generic.py.)"""
    ref = rh.load_reference()  # (puts the shim of the un-vendored arm_pytorch_utilities on the path)
    import arm_pytorch_utilities.model.make as make
    torch.manual_seed(1)
    ctrl, ds, env = rh.build_task('push', num_samples=80, rollout_samples=2, horizon=10)
    ctrl.set_goal(np.array([0.55, -0.65, 0., 0., 0.]))
    nm = ctrl.dynamics.nominal_model
    inv = nm.ds.preprocessor.tsf.transforms[1].tsf

    def net(n_in, n_out, h_units, out_scale):
        cfg = ref.load_data.DataConfig()
        cfg.nx, cfg.ny = n_in, n_out
        m = make.make_sequential_network(cfg, h_units=h_units).double()
        with torch.no_grad():
            list(m.modules())[-1].weight.mul_(out_scale)  # keep the untrained rollout bounded
        return m

    nz, nv, ne = 5, int(inv.nv), inv.x_extractor.out_features
    inv.encoder.model = net(8, nz, enc_h, 1.0)
    inv.extracted_linear_decoder.model = net(ne, 5 * nv, dec_h, 0.2)
    nm.model = nm.user.model = net(nz, nv, dyn_h, 1.0)  # (NetworkModelWrapper keeps the network under both names, model.py:212-213)
    spec = extract.spec_from_controller(ctrl)
    d = spec.dynamics
    assert d.encoder.dims == [8] + list(enc_h) + [nz] and d.net.dims == [nz] + list(dyn_h) + [nv]
    assert d.decoder.dims == [ne] + list(dec_h) + [5 * nv]
    g = torch.Generator().manual_seed(5)
    state = torch.randn(5, dtype=torch.double, generator=g) * 0.4
    U0 = torch.rand(spec.sampler.T, 3, dtype=torch.double, generator=g) * 0.5
    noise = O.sample_noise(spec, 9)
    a_ref = _run_ref(ctrl, spec, state, U0, noise)
    out = O.command(spec, state, U0, noise)
    rel = ((out['cost_total'] - ctrl.mpc.cost_total).abs() / ctrl.mpc.cost_total.abs()).max().item()
    assert rel < 1e-12 and out['argmin'] == int(ctrl.mpc.cost_total.argmin())
    assert (out['action'] - a_ref).abs().max().item() < 1e-12
