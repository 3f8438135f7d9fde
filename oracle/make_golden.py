"""Generate tests/golden/*.npz by running the reference's UNMODIFIED in-tree MPPI path (over oracle/shim).

Run in the container that has /root/reference:   python oracle/make_golden.py
Each fixture holds the extracted RolloutSpec (tampc_b200.spec.spec_to_arrays), the inputs (state, nominal U
before the shift, injected noise) and the reference's outputs (cost_total, action, updated U, argmin, omega,
a few rollout states, get_rollouts).  The GPU box has no /root/reference: tests there read these files.
TEST INFRASTRUCTURE ONLY."""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import ref_harness as rh  # noqa: E402
from tampc_b200 import extract, spec as S  # noqa: E402
from tampc_b200 import analytic  # noqa: E402
from oracle import mppi_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

# fixtures lifted from the reference scripts
PUSH_TRAP = ([0.6147, 0.1381, -1.2658, 6.9630, 14.9701], [-0.5439, 0.6192, -0.5857])  # push_main.py:964-965
PUSH_GOAL = [0.55, -0.65, 0., 0., 0.]  # level 5, push_main.py:138-141
PUSH_START = [-0.4, 0.23, -math.pi / 5.2, 0., 0.]
PEG_TRAPS = [([-3.5530e-03, 1.4122e-01, 1.9547e-01, 7.1541e-01, -1.0235e+01], [0.3596, 0.2701]),
             ([-0.0775, 0.1415, 0.1956, -9.7136, -10.1923], [0.2633, 0.3216]),
             ([0.1663, 0.1417, 0.1956, 4.4845, -10.2436], [1.0000, 0.8479]),
             ([0.0384, 0.1412, 0.1955, 4.6210, -10.4069], [0.3978, 0.6424])]  # peg_main.py:431-442
PEG_TRAP_WEIGHT = 0.0035  # peg_main.py:443
PEG_START = [-3.35635743e-03, 3.24323310e-01, 1.95463138e-01, 5.71478642e+00, -4.01359529e+00]  # peg_main.py:444
PEG_GOAL = [0., 0.2, 0.1955, 0., 0.]  # hole (0,0.2) level 5 (peg_main.py:118-120), z of the fixtures


def t64(a):
    return torch.tensor(a, dtype=torch.double)


def run_reference_command(ctrl, spec, state, U0, noise, gp_eps=None):
    """mpc.command(state) of the reference with the noise draw injected (and, for a GP-sampled model, the standard
    normals behind OnlineGPMixing.sample: gp_eps (T, M, K, ny), one (M*K, ny) block per horizon step)."""
    mpc = ctrl.mpc
    mpc.U = U0.clone()
    orig = mpc.noise_dist.sample
    KT = (spec.sampler.K, spec.sampler.T)
    mpc.noise_dist.sample = lambda shape=(): noise.clone() if tuple(shape) == KT else orig(shape)
    draws = [] if gp_eps is None else [gp_eps[t].reshape(-1, gp_eps.shape[-1]) for t in range(gp_eps.shape[0])]
    try:
        with torch.no_grad(), rh.inject_gp_draws(draws) as inj:
            action = mpc.command(state.clone())
            assert not inj.draws, "the reference consumed {} of the injected GP draws".format(len(draws) - len(inj.draws))
    finally:
        mpc.noise_dist.sample = orig
    return action


def save_case(name, ctrl, state, seed, note):
    spec = extract.spec_from_controller(ctrl, name)
    mpc = ctrl.mpc
    g = torch.Generator().manual_seed(seed + 1)
    # nominal sequence before the shift: a plausible warm start inside the bounds
    U0 = t64(spec.sampler.noise_mu) + 0.3 * torch.randn(spec.sampler.T, spec.sampler.nu, dtype=torch.double, generator=g)
    if spec.sampler.u_min is not None:
        U0 = torch.max(torch.min(U0, t64(spec.sampler.u_max)), t64(spec.sampler.u_min))
    noise = O.sample_noise(spec, seed)
    state = t64(state)
    gp_eps = None
    if spec.dynamics.gp is not None:
        ny = spec.dynamics.gp.resid.shape[1]
        gp_eps = torch.randn(spec.sampler.T, mpc.M, spec.sampler.K, ny, dtype=torch.double,
                             generator=torch.Generator().manual_seed(seed + 2))
    action = run_reference_command(ctrl, spec, state, U0, noise, gp_eps)
    with torch.no_grad(), rh.inject_gp_draws([]):  # a GP-sampled model rolls the nominal sequence out at its mean
        rollouts = mpc.get_rollouts(state.clone())[0]
    M = mpc.M
    states = mpc.states  # (M,K,T,nx)
    ref = {
        'in.state': state.numpy(), 'in.U': U0.numpy(), 'in.noise': noise.numpy(), 'in.noise_seed': np.array(seed),
        'ref.cost_total': mpc.cost_total.numpy(), 'ref.action': action.numpy(), 'ref.U': mpc.U.numpy(),
        'ref.argmin': np.array(int(torch.argmin(mpc.cost_total))), 'ref.omega': mpc.omega.numpy(),
        'ref.states_first8': states[0, :8].numpy(), 'ref.get_rollouts': rollouts.numpy(),
        'ref.M': np.array(M), 'note': np.array(note),
    }
    if gp_eps is not None:
        ref['in.gp_eps'] = gp_eps.numpy()
    # oracle must agree before the fixture is written
    out = O.command(spec, state, U0, noise, gp_eps)
    rel = ((out['cost_total'] - mpc.cost_total).abs() / mpc.cost_total.abs().clamp_min(1e-300)).max().item()
    aerr = (out['action'] - action).abs().max().item()
    assert out['argmin'] == int(ref['ref.argmin']), (name, out['argmin'], int(ref['ref.argmin']))
    assert rel < 1e-12 and aerr < 1e-10, (name, rel, aerr)
    arrays = dict(S.spec_to_arrays(spec))
    arrays.update(ref)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('{:28s} K={:4d} T={:2d} M={:2d} cost[{:.4g},{:.4g}] oracle rel {:.1e} act {:.1e}  {:.0f} KB'.format(
        name, spec.sampler.K, spec.sampler.T, M, float(mpc.cost_total.min()), float(mpc.cost_total.max()), rel, aerr,
        os.path.getsize(path) / 1024))
    return spec


def add_push_traps(ctrl, n=3):
    tx, tu = t64(PUSH_TRAP[0]), t64(PUSH_TRAP[1])
    ctrl.trap_set.append((tx, tu))
    for dx in (-0.1, 0.1)[:n - 1]:  # +2 synthetic copies shifted by +-0.1 in x
        t2 = tx.clone()
        t2[0] += dx
        ctrl.trap_set.append((t2, tu.clone()))


def gp_cases():
    """the online GP residual (OnlineGPMixing over oracle/shim/gpytorch), both placements."""
    # 9. --always_estimate_error (push_main.py:1788-1789): the nominal model IS a GP around the latent network,
    #    built on the nominal datasource (first 50 training rows, hybrid_model.py:110-112), script M=10
    ctrl, ds, env = rh.build_task('push', num_samples=32, rollout_samples=10, horizon=12,
                                  nominal_adapt='GP_KERNEL_INDEP_OUT')
    ctrl.set_goal(np.array(PUSH_GOAL))
    add_push_traps(ctrl, n=1)
    save_case('push_gp_latent', ctrl, PUSH_START, 41, 'GP residual in the latent space (nominal_adapt=GP_KERNEL_INDEP_OUT), M=10')

    # 10. the TAMPC flow: entering non-nominal dynamics swaps in a temporary local GP with no data, behind the plain
    #     scaler preprocessor, and feeds it the transitions seen since (online_controller.py:324-335; hybrid_model.py:
    #     189-199).  Here: six "stuck against the wall" transitions.  rollout_var_cost at the class default 0.5
    #     (controller.py:317) so that the across-replica variance term is exercised.
    ctrl, ds, env = rh.build_task('push', num_samples=32, rollout_samples=10, horizon=12)
    ctrl.set_goal(np.array(PUSH_GOAL))
    add_push_traps(ctrl, n=1)
    ctrl.mpc.rollout_var_cost = 0.5
    g = torch.Generator().manual_seed(8)
    x = t64(PUSH_TRAP[0])
    with torch.enable_grad():
        ctrl.dynamics.use_temp_local_nominal_model()
        for i in range(6):
            u = t64(spec_bounds_sample(ctrl, g))
            nx_ = x + 0.003 * torch.randn(5, dtype=torch.double, generator=g) * t64([1, 1, 1, 50, 50])
            ctrl.dynamics.update(x, u, nx_)
            x = nx_
    save_case('push_gp_local', ctrl, x.tolist(), 43, 'temporary local GP (6 points, original space), M=10, var cost 0.5')


def gp_fit_case():
    """the hyper-parameter trajectory of the reference's OnlineGPMixing (Adam through the restated
    gpytorch, oracle/shim/gpytorch) — the constructor's 150 iterations on 30 nominal points, then four `update`s of 15
    iterations each (optimiser state carried over, online_model.py:386-430).  The device fit (tampc_gp_fit) must land
    on the same raw parameters from the same (inputs, targets - mean function) at every stage."""
    ref = rh.load_reference()
    ctrl, ds, env = rh.build_task('push', num_samples=8, rollout_samples=2, horizon=4)
    with torch.enable_grad():
        om = ref.hybrid_model.HybridDynamicsModel.get_local_model(
            env.state_difference, ctrl.dynamics.pm, 'cpu', ds, preprocessor=ref.util.no_tsf_preprocessor(), allow_update=True,
            train_slice=slice(0, 30))
    arrays = {}

    def record(stage, iters):
        with torch.no_grad():
            mean = om.gp.forward(om.xu).mean
        arrays['stage{}.X'.format(stage)] = om.xu.detach().numpy().copy()
        arrays['stage{}.resid'.format(stage)] = (om.y - mean).detach().numpy().copy()
        arrays['stage{}.iters'.format(stage)] = np.array(iters)
        arrays['stage{}.raw_lengthscale'.format(stage)] = om.gp.covar_module.base_kernel.raw_lengthscale.detach().numpy().reshape(-1).copy()
        arrays['stage{}.raw_outputscale'.format(stage)] = om.gp.covar_module.raw_outputscale.detach().numpy().reshape(-1).copy()
        arrays['stage{}.raw_task_noises'.format(stage)] = om.likelihood.raw_task_noises.detach().numpy().reshape(-1).copy()
        arrays['stage{}.raw_noise'.format(stage)] = om.likelihood.raw_noise.detach().numpy().reshape(-1).copy()
        arrays['stage{}.last_loss'.format(stage)] = np.array(om.last_loss)

    record(0, om.training_iter)
    g = torch.Generator().manual_seed(12)
    x = t64(PUSH_TRAP[0])
    for stage in range(1, 5):
        u = t64(spec_bounds_sample(ctrl, g))
        nx_ = x + 0.003 * torch.randn(5, dtype=torch.double, generator=g) * t64([1, 1, 1, 50, 50])
        with torch.enable_grad():
            om.update(x, u, nx_)
        x = nx_
        record(stage, om.training_iter // 10)
    arrays['n_stages'] = np.array(5)
    path = os.path.join(OUT, 'gp_fit.npz')
    np.savez_compressed(path, **arrays)
    print('gp_fit: N {} -> {}, final loss {:.6f}, lengthscale raw {}  {:.0f} KB'.format(
        len(arrays['stage0.X']), len(arrays['stage4.X']), float(arrays['stage4.last_loss']), arrays['stage4.raw_lengthscale'].round(3),
        os.path.getsize(path) / 1024))


def spec_bounds_sample(ctrl, g):
    lo, hi = ctrl.u_min.double(), ctrl.u_max.double()
    return (lo + (hi - lo) * torch.rand(lo.shape[0], dtype=torch.double, generator=g)).tolist()


def main():
    torch.manual_seed(0)
    np.random.seed(0)
    if '--gp-only' in sys.argv:
        return gp_cases()
    if '--gp-fit-only' in sys.argv:
        return gp_fit_case()

    # 1. pushing, nominal cost (QR + 3 traps + terminal x50), M=2
    ctrl, ds, env = rh.build_task('push', num_samples=256, rollout_samples=2)
    ctrl.set_goal(np.array(PUSH_GOAL))
    add_push_traps(ctrl)
    save_case('push_nominal', ctrl, PUSH_START, 1234, 'push REX_EXTRACT, QR+3 traps+terminal, M=2')

    # 2. same at the script's M=10: deterministic replicas change nothing
    ctrl, ds, env = rh.build_task('push', num_samples=64, rollout_samples=10)
    ctrl.set_goal(np.array(PUSH_GOAL))
    add_push_traps(ctrl)
    ctrl.trap_set_weight = 0.37
    save_case('push_M10', ctrl, [0.4, 0.15, -math.pi / 8, 0, 0], 77, 'push at the script M=10, state near the trap')

    # 3. recovery mode: MAB-weighted pair of GoalSetCosts, horizon 5, no terminal (online_controller.py:337-402)
    ctrl, ds, env = rh.build_task('push', num_samples=128, rollout_samples=2)
    ctrl.set_goal(np.array(PUSH_GOAL))
    ctrl.nominal_max_velocity = 0.012  # push_main.py:979-989 tune_recovery_policy fixture
    add_push_traps(ctrl, n=1)
    g = torch.Generator().manual_seed(5)
    base = t64([-0.0525, 0.1609, -0.4940, -3.0287, 15.0308])
    hist = [base + 0.02 * i * t64([1, 0.5, 0.3, 0, 0]) + 0.01 * torch.randn(5, dtype=torch.double, generator=g) for i in range(9)]
    ctrl.x_history = list(hist)
    ctrl.u_history = [t64([0.1, 0.5, 0.1]) for _ in hist]
    ctrl.nominal_dynamic_states = [list(hist[:6])]
    ctrl.nonnominal_dynamics_start_index = 3
    ctrl.single_step_move_dist = [(i, 0.001 * ((7 * i) % 5 + 1), 0.01) for i in range(len(hist) - 1)]
    ctrl._start_recovery_mode()
    ctrl.last_arm_pulled = torch.tensor(7)
    assert ctrl.mpc.T == 5 and ctrl.mpc.terminal_state_cost is None
    save_case('push_recovery', ctrl, [0.063, 0.14, -1.246, 0., 0.], 99, 'recovery cost: 2 goal sets, MAB arm 7, T=5')

    # 4. bad-state guard: ||state|| > 1e5 at t=0 (online_controller.py:63-66,77)
    ctrl, ds, env = rh.build_task('push', num_samples=64, rollout_samples=2, horizon=12)
    ctrl.set_goal(np.array(PUSH_GOAL))
    add_push_traps(ctrl, n=1)
    save_case('push_badstate', ctrl, [2e5, 0.1, 0.3, 0., 0.], 11, 'start state with norm > 1e5: zeroed state/action quirk')

    # 5. sample_null_action (controller.py:389-390)
    ctrl, ds, env = rh.build_task('push', num_samples=64, rollout_samples=2, horizon=10)
    ctrl.mpc.sample_null_action = True
    ctrl.set_goal(np.array(PUSH_GOAL))
    save_case('push_null_action', ctrl, PUSH_START, 21, 'sample_null_action=True, empty trap set')

    # 6. peg, 4 traps at weight 0.0035
    ctrl, ds, env = rh.build_task('peg', num_samples=256, rollout_samples=2)
    ctrl.set_goal(np.array(PEG_GOAL))
    for x, u in PEG_TRAPS:
        ctrl.trap_set.append((t64(x), t64(u)))
    ctrl.trap_set_weight = torch.tensor([PEG_TRAP_WEIGHT], dtype=torch.double)
    save_case('peg_nominal', ctrl, PEG_START, 4321, 'peg REX_EXTRACT, QR+4 traps (w=0.0035)+terminal, M=2')

    # 7. no-transform network: PytorchTransformer(RobustMinMaxScaler) both sides, random-init 8->32->32->5
    ref = rh.load_reference()
    from arm_pytorch_utilities.model import make
    ctrl, ds, env = rh.build_task('push', num_samples=128, rollout_samples=2, horizon=20)
    ds.update_preprocessor(ref.util.no_tsf_preprocessor())
    torch.manual_seed(3)
    mw = ref.util.TranslationNetworkWrapper(
        ref.model.DeterministicUser(make.make_sequential_network(ds.config)), ds, name='golden_notransform')
    pm = ref.prior.NNPrior(mw)
    hybrid = ref.hybrid_model.HybridDynamicsModel([ds], pm, env.state_difference, ['NO_TRANSFORM'], device='cpu',
                                                  preprocessor=ref.util.no_tsf_preprocessor(),
                                                  nominal_model_kwargs={'online_adapt': ref.hybrid_model.OnlineAdapt.NONE})
    common, mpc_opts = rh.push_options(ref)
    mpc_opts.update(num_samples=128, rollout_samples=2, horizon=20)
    with torch.no_grad():
        ctrl = ref.online_controller.OnlineMPPI(ds, hybrid, ds.original_config(),
                                                gating=ref.gating_function.AlwaysSelectNominal(),
                                                constrain_state=rh.push_constrain_state, mpc_opts=mpc_opts, **common)
    ctrl.set_goal(np.array(PUSH_GOAL))
    add_push_traps(ctrl, n=2)
    save_case('push_mlp_notransform', ctrl, PUSH_START, 31, 'plain MLP with in/out RobustMinMax scalers (random init)')

    # 8. pendulum with analytic dynamics, script defaults K=100,T=15,M=20,lambda=1 (pendulum.py:111-123,262-272)
    pend = analytic.Pendulum()

    class _DS:  # MPC.__init__ only asks for a training set to log the model error (controller.py:206-214)
        def training_set(self, original=False):
            g = torch.Generator().manual_seed(0)
            xu = torch.rand(16, 3, dtype=torch.double, generator=g)
            return xu, torch.zeros(16, 2, dtype=torch.double), None

    cfg = ref.load_data.DataConfig(predict_difference=True, predict_all_dims=True)
    cfg.nx, cfg.nu, cfg.ny = 2, 1, 2

    def pend_compare_to_goal(state, goal):  # pendulum.py:66-71
        if goal.dim() == 1:
            goal = goal.view(1, -1)
        dtheta = ref.math_utils.angular_diff_batch(state[:, 0], goal[:, 0])
        return torch.cat((dtheta.view(-1, 1), (state[:, 1] - goal[:, 1]).view(-1, 1)), dim=1)

    ctrl = ref.controller.MPPI_MPC(_DS(), pend, cfg, Q=t64([[1, 0], [0, 0.1]]), R=0.001, u_max=2.0, u_min=-2.0,
                                   compare_to_goal=pend_compare_to_goal, device='cpu', use_trap_cost=False,
                                   mpc_opts={'num_samples': 100, 'horizon': 15, 'lambda_': 1,
                                             'noise_sigma': torch.eye(1, dtype=torch.double)})
    ctrl.set_goal(np.array([0., 0.]))
    save_case('pendulum', ctrl, [math.pi, 1.0], 6, 'analytic pendulum, K=100 T=15 M=20 lambda=1, no trap cost')

    gp_cases()
    gp_fit_case()


if __name__ == '__main__':
    main()
