"""GPU: the integration path a reference user takes — install() on a live controller — exercised with a duck-typed
controller rebuilt from the golden specs (tests/fake_tampc.py); outputs must match the reference's golden outputs."""
import numpy as np
import pytest
import torch

from conftest import load_golden
import fake_tampc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', ['push_nominal', 'peg_nominal', 'push_recovery', 'push_mlp_notransform', 'push_null_action'])
def test_install_on_live_controller_reproduces_reference(cuda_lib, case):
    from tampc_b200 import extract, spec as S
    from tampc_b200.mppi import install
    z, spec = load_golden(case)
    ctrl = fake_tampc.make_controller(spec, z['in.U'])
    got = extract.spec_from_controller(ctrl)
    a, b = S.spec_to_arrays(spec), S.spec_to_arrays(got)
    for k in a:
        if k.endswith('name') or 'bad_state_norm' in k:
            continue
        np.testing.assert_allclose(np.asarray(a[k], dtype=float), np.asarray(b[k], dtype=float), rtol=0, atol=1e-12, err_msg=k)
    mpc = install(ctrl, precision='f64')
    assert ctrl.mpc is mpc
    # OnlineMPC.predict_next_state on the device: the nominal network only (online_controller.py:47-58)
    from oracle import mppi_oracle as O
    x, u = torch.from_numpy(z['in.state']), torch.from_numpy(z['in.U'][0])
    want = O.predict(spec.dynamics, x.view(1, -1), u.view(1, -1)).numpy()
    np.testing.assert_allclose(ctrl.predict_next_state(x, u), want, rtol=0, atol=1e-10)
    action = ctrl.mpc.command(torch.from_numpy(z['in.state']), noise=z['in.noise']).numpy()
    np.testing.assert_allclose(action, z['ref.action'], rtol=0, atol=1e-8)
    rel = np.max(np.abs(mpc.cost_total.numpy() - z['ref.cost_total']) / np.abs(z['ref.cost_total']))
    assert rel < 1e-9 and mpc.stats['argmin'] == int(z['ref.argmin'])
    mpc.close()


def test_cost_program_is_reread_every_command(cuda_lib):
    """The controller mutates trap set / weights / cost callbacks between commands (online_controller.py:353,
    400-402,416-418,461-478); the drop-in must follow without being told."""
    from tampc_b200.mppi import install
    from oracle import mppi_oracle as O
    from tampc_b200 import spec as S
    z, spec = load_golden('push_nominal')
    zr, spec_r = load_golden('push_recovery')
    ctrl = fake_tampc.make_controller(spec, z['in.U'])
    mpc = install(ctrl, precision='f64')
    state = torch.from_numpy(z['in.state'])
    # 1. anneal the trap weight and grow the trap set
    ctrl.trap_set_weight = 0.37
    ctrl.trap_set.append((torch.tensor([0.1, 0.2, 0.3, 0., 0.], dtype=torch.double), torch.tensor([0.5, 0.5, -0.5], dtype=torch.double)))
    noise = O.sample_noise(spec, 3)
    U0 = mpc.U.clone()
    mpc.command(state, noise=noise.numpy())
    spec.cost.trap_weight = 0.37
    spec.cost.trap_x = np.vstack([spec.cost.trap_x, [0.1, 0.2, 0.3, 0., 0.]])
    spec.cost.trap_u = np.vstack([spec.cost.trap_u, [0.5, 0.5, -0.5]])
    want = O.command(spec, state, U0, noise)
    assert float(((mpc.cost_total - want['cost_total']).abs() / want['cost_total'].abs()).max()) < 1e-9
    # 2. enter recovery mode the way OnlineMPPI._start_recovery_mode does
    rc = fake_tampc.make_controller(spec_r, zr['in.U'])
    ctrl.recovery_cost = rc.recovery_cost
    ctrl.mpc.running_cost = ctrl._recovery_running_cost
    ctrl.mpc.terminal_state_cost = None
    ctrl.mpc.change_horizon(5)
    U1 = mpc.U.clone()
    noise5 = O.sample_noise(spec_r, 4, K=spec.sampler.K, T=5)
    mpc.command(state, noise=noise5.numpy())
    spec_mix = S.RolloutSpec(spec.dynamics, spec_r.cost, S.SamplerSpec(K=spec.sampler.K, T=5, nu=3, lambda_=spec.sampler.lambda_,
                             noise_mu=spec.sampler.noise_mu, noise_sigma=spec.sampler.noise_sigma, u_min=spec.sampler.u_min,
                             u_max=spec.sampler.u_max, u_init=spec.sampler.u_init))
    want = O.command(spec_mix, state, U1, noise5)
    assert float(((mpc.cost_total - want['cost_total']).abs() / want['cost_total'].abs()).max()) < 1e-9
    # 3. leave it again
    ctrl.mpc.running_cost = ctrl._running_cost
    ctrl.mpc.terminal_state_cost = ctrl._terminal_cost
    ctrl.mpc.change_horizon(40)
    mpc.command(state)
    # 4. an unknown callback is an error, not a fallback
    from tampc_b200.extract import UnsupportedSpec
    ctrl.mpc.running_cost = lambda x, u: x.sum(dim=1)
    with pytest.raises(UnsupportedSpec):
        mpc.command(state)
    mpc.close()


@pytest.mark.parametrize('case', ['push_gp_local', 'push_gp_latent'])
def test_online_gp_is_followed_when_the_nominal_model_is_swapped(cuda_lib, case):
    """online_controller.py:324-335 swaps `dynamics.nominal_model` for an OnlineGPMixing when non-nominal dynamics
    start and back when they end (hybrid_model.py:189-202); every dynamics.update refits it.  The drop-in re-reads the
    residual on each command."""
    from tampc_b200.mppi import install
    z, spec = load_golden(case)
    gp_spec = spec.dynamics.gp
    spec.dynamics.gp = None
    ctrl = fake_tampc.make_controller(spec, z['in.U'])
    mpc = install(ctrl, precision='f64')
    nominal = ctrl.dynamics.nominal_model
    mpc.command(torch.from_numpy(z['in.state']), noise=z['in.noise'])
    cost_nominal = mpc.cost_total.numpy().copy()
    # entering non-nominal dynamics
    spec.dynamics.gp = gp_spec
    ctrl.dynamics.nominal_model = fake_tampc.make_online_gp(spec, nominal)
    mpc.U = torch.from_numpy(z['in.U'])
    action = mpc.command(torch.from_numpy(z['in.state']), noise=z['in.noise'], gp_eps=z['in.gp_eps']).numpy()
    rel = np.max(np.abs(mpc.cost_total.numpy() - z['ref.cost_total']) / np.abs(z['ref.cost_total']))
    assert rel < 1e-9 and mpc.stats['argmin'] == int(z['ref.argmin'])
    np.testing.assert_allclose(action, z['ref.action'], rtol=0, atol=1e-8)
    # predict_next_state stays the ORIGINAL nominal network while the GP serves the rollouts (online_controller.py:47-58):
    # its error against the observed state is what detects (leaving) non-nominal dynamics
    from oracle import mppi_oracle as O
    import copy
    d_nom = copy.copy(spec.dynamics)
    d_nom.gp = None
    x, u = torch.from_numpy(z['in.state']), torch.from_numpy(z['in.U'][0])
    want = O.predict(d_nom, x.view(1, -1), u.view(1, -1)).numpy()
    np.testing.assert_allclose(ctrl.predict_next_state(x, u), want, rtol=0, atol=1e-10)
    big = torch.tensor([2e5, 0., 9., 0., 0.], dtype=torch.double)  # beyond the guard, yaw outside (-pi, pi]: neither applies here
    np.testing.assert_allclose(ctrl.predict_next_state(big, u), O.predict(d_nom, big.view(1, -1), u.view(1, -1)).numpy(), rtol=1e-9, atol=1e-6)
    assert np.max(np.abs(mpc.predict_next_state(x, u) - want)) > 1e-9  # (the full _apply_dynamics step does include the residual)
    # a GP update between commands (new data point): the next command must see it
    g2 = fake_tampc.make_online_gp(spec, nominal)
    g2.xu, g2.y = g2.xu[:-1], g2.y[:-1]
    ctrl.dynamics.nominal_model.xu, ctrl.dynamics.nominal_model.y = g2.xu, g2.y
    mpc.U = torch.from_numpy(z['in.U'])
    mpc.command(torch.from_numpy(z['in.state']), noise=z['in.noise'], gp_eps=z['in.gp_eps'])
    assert np.max(np.abs(mpc.cost_total.numpy() - z['ref.cost_total'])) > 0
    # leaving: back to the nominal network (and to the tensor-core layout)
    ctrl.dynamics.nominal_model = nominal
    mpc.U = torch.from_numpy(z['in.U'])
    mpc.command(torch.from_numpy(z['in.state']), noise=z['in.noise'])
    np.testing.assert_array_equal(mpc.cost_total.numpy(), cost_nominal)
    mpc.close()


def test_packaged_spec_starts_a_controller_without_the_datasets(cuda_lib, tmp_path):
    """B200MPPI.from_package reproduces the reference outputs from the .npz alone."""
    from tampc_b200 import spec as S
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden('peg_nominal')
    path = str(tmp_path / 'peg.spec.npz')
    S.save_spec(path, spec)
    m = B200MPPI.from_package(path, precision='f64', U_init=torch.from_numpy(z['in.U']))
    action = m.command(z['in.state'], noise=z['in.noise']).numpy()
    np.testing.assert_allclose(action, z['ref.action'], rtol=0, atol=1e-8)
    assert m.stats['argmin'] == int(z['ref.argmin'])
    m.close()


@pytest.mark.parametrize('case,precision', [('push_nominal', 'f64'), ('peg_nominal', 'tf32x3'), ('push_gp_local', 'f64')])
def test_one_step_sampler_of_the_apf_baselines(cuda_lib, case, precision):
    """APFVO / APFSP push 5000 uniformly sampled actions through _apply_dynamics from the current state and take the
    argmin of a potential (online_controller.py:583-590, 657-664): the dynamics step runs in the fused kernel."""
    from oracle import mppi_oracle as O
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden(case)
    if spec.dynamics.gp is not None:
        spec.dynamics.gp.sample = False  # the posterior mean: comparable without injecting draws
        spec.sampler.rollout_samples = 1
    ctrl = fake_tampc.make_controller(spec, z['in.U'])
    if spec.dynamics.gp is not None:
        ctrl.dynamics.nominal_model = fake_tampc.make_online_gp(spec, ctrl.dynamics.nominal_model)
    ctrl.u_min, ctrl.u_max = torch.from_numpy(spec.sampler.u_min), torch.from_numpy(spec.sampler.u_max)
    m = B200MPPI.for_one_step(ctrl, samples=5000, precision=precision)
    g = torch.Generator().manual_seed(0)
    u = torch.rand(5000, spec.dynamics.nu, dtype=torch.double, generator=g) * (ctrl.u_max - ctrl.u_min) + ctrl.u_min
    x = torch.from_numpy(z['in.state'])
    got = m.apply_dynamics(x, u)
    want, _ = O.apply_dynamics(spec.dynamics, x.view(1, -1).repeat(5000, 1), u)
    # float32 networks: the reaction-force dims are O(10) and un-scaled by 1/0.12 (float32 state ulp there is 1e-6)
    tol = 1e-10 if precision == 'f64' else 1e-4
    np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=tol, atol=tol)
    # the sampler's decision: argmin of a potential over the next states (here the goal cost, cost.py:11-18)
    goal = torch.from_numpy(spec.cost.goal if spec.cost.kind == 0 else np.zeros(spec.dynamics.nx))
    pot = lambda s: ((s - goal)[:, :2] ** 2).sum(dim=1)
    assert int(pot(got).argmin()) == int(pot(want).argmin())
    # fewer actions than the handle's K, and a bad start state (guard: rows come back as zeros)
    got7 = m.apply_dynamics(x, u[:7])
    np.testing.assert_allclose(got7.numpy(), want[:7].numpy(), rtol=tol, atol=tol)
    if spec.dynamics.bad_state_norm is not None:
        bad = m.apply_dynamics(torch.full((spec.dynamics.nx,), 1e6, dtype=torch.double), u[:3])
        assert torch.all(bad == 0)
    m.close()


@pytest.mark.parametrize('case,precision', [('push_nominal', 'f64'), ('push_nominal', 'tf32x3'), ('peg_nominal', 'mixed')])
def test_apf_action_selection_on_the_device(cuda_lib, case, precision):
    """APFVO / APFSP (online_controller.py:583-590, 657-664) — dynamics, potential and argmin in one
    device call — against the oracle (which tests/test_reference_live.py pins to the reference's cost classes)."""
    import types
    from oracle import mppi_oracle as O
    from tampc_b200 import spec as S
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden(case)
    ctrl = fake_tampc.make_controller(spec, z['in.U'])
    ctrl.u_min, ctrl.u_max = torch.from_numpy(spec.sampler.u_min), torch.from_numpy(spec.sampler.u_max)
    nx, nu = spec.dynamics.nx, spec.dynamics.nu
    g = torch.Generator().manual_seed(2)
    x = torch.from_numpy(z['in.state'])
    u = torch.rand(5000, nu, dtype=torch.double, generator=g) * (ctrl.u_max - ctrl.u_min) + ctrl.u_min
    nxt0, _ = O.apply_dynamics(spec.dynamics, x.view(1, -1).repeat(16, 1), u[:16])
    # APFVO: goal + artificial repulsion; repulsive states where the model thinks some actions lead (online_controller.py:578-581)
    rep = types.SimpleNamespace(state_set=[nxt0[3].clone(), nxt0[11].clone() + 0.05, x.clone() + 0.2], max_dist_influence=1.0, gain=1.0,
                                compare_to_goal=ctrl.compare_to_goal)
    ctrl.cost = types.SimpleNamespace(fn=[ctrl.goal_cost, rep], weights=None)
    m = B200MPPI.for_one_step(ctrl, samples=5000, this_cost=None, precision=precision)
    tol = 1e-9 if precision == 'f64' else 1e-5

    def check(n):
        cost_spec = m._cost_spec
        idx, a, c, nxt = m.select_action(x, u[:n])
        wi, wc, wn = O.select_action(S.RolloutSpec(spec.dynamics, cost_spec, m.spec.sampler), x, u[:n])
        s = torch.sort(wc).values
        gap = float((s[1] - s[0]) / s[0].abs().clamp_min(1e-300)) if n > 1 else 1.0
        # the potential of the state the DEVICE reached (a repulsive 1/rho^2 term amplifies the float32 state error of
        # the network step, which is checked separately below): float32 cost arithmetic only
        own = float(O.apf_cost(cost_spec, nxt.view(1, -1))[0])
        print(case, precision, 'n', n, 'top-2 gap', gap, 'cost err vs oracle dynamics', abs(c - float(wc[idx])) / abs(float(wc[idx])),
              'vs oracle potential of the device state', abs(c - own) / abs(own))
        assert abs(c - own) <= tol * max(1.0, abs(own))
        assert abs(c - float(wc[idx])) <= 1e-4 * max(1.0, abs(float(wc[idx])))
        if gap > 1e-4:
            assert idx == wi and torch.equal(a, u[idx])
        else:
            print('   argmin assert skipped: top-2 gap', gap)
        np.testing.assert_allclose(nxt.numpy(), wn[idx].numpy(), rtol=1e-4 if precision != 'f64' else 1e-10, atol=1e-4 if precision != 'f64' else 1e-10)

    check(5000)
    check(37)
    check(1)
    # the trap set grows, then APFSP's switched potentials: helicoid around an obstacle / plain goal potential
    rep.state_set.append(nxt0[5].clone())
    m.set_apf_cost()
    assert len(m._cost_spec.trap_x) == 4
    check(5000)
    hel = types.SimpleNamespace(rxy=x[:2].clone(), oxy=x[:2].clone() + torch.tensor([0.15, -0.1], dtype=torch.double), clockwise=False)
    m.set_apf_cost(hel)
    check(5000)
    hel.clockwise = True
    m.set_apf_cost(hel)
    check(4999)
    m.set_apf_cost(ctrl.goal_cost)
    check(5000)
    m.close()


def test_trapset_cost_normalisation_on_the_device(cuda_lib):
    """OnlineMPPI.normalize_trapset_cost_to_state (online_controller.py:517-530) served from the installed cost program."""
    from oracle import mppi_oracle as O
    from tampc_b200.mppi import install
    z, spec = load_golden('push_nominal')
    ctrl = fake_tampc.make_controller(spec, z['in.U'])
    ctrl.normalize_trapset_cost_to_state = lambda x: None
    mpc = install(ctrl, precision='tf32x3')
    assert ctrl.normalize_trapset_cost_to_state == mpc.normalize_trapset_cost_to_state
    g = torch.Generator().manual_seed(9)
    for i in range(4):
        x = torch.randn(5, dtype=torch.double, generator=g) * 0.5
        want, _, _ = O.normalize_trapset_cost_to_state(spec.cost, x)
        got = ctrl.normalize_trapset_cost_to_state(x)
        assert got.shape == (1,) and abs(float(got) - float(want)) <= 1e-12 * abs(float(want))
    # the trap set changed since the last command: the device program follows; no traps -> None like the reference
    ctrl.trap_set.append((torch.tensor([0.3, -0.1, 0.4, 0., 0.], dtype=torch.double), torch.tensor([0.1, 0.9, 0.2], dtype=torch.double)))
    spec.cost.trap_x = np.vstack([spec.cost.trap_x, [0.3, -0.1, 0.4, 0., 0.]])
    x = torch.tensor([0.2, 0.1, -0.3, 1., 2.], dtype=torch.double)
    want, _, _ = O.normalize_trapset_cost_to_state(spec.cost, x)
    assert abs(float(ctrl.normalize_trapset_cost_to_state(x)) - float(want)) <= 1e-12 * abs(float(want))
    ctrl.trap_set.clear()
    assert ctrl.normalize_trapset_cost_to_state(x) is None
    mpc.close()
