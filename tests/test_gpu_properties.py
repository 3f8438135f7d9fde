"""GPU: size-independent properties at BASELINE.json's full sizes, oracle parity at sizes the oracle finishes in
seconds, the on-device Philox sampler, and the rest of the MPPI surface (horizon changes, rollouts, error paths)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mppi_oracle as O
from tampc_b200 import spec as S

pytestmark = pytest.mark.gpu

PRECISIONS = ['f64', 'tf32x3', 'f16x3', 'mixed']
RTOL = {'f64': 1e-9, 'tf32x3': 1e-5, 'f16x3': 1e-5, 'mixed': 1e-5, 'f32': 1e-5}


def _spec(case, K=None, T=None):
    z, spec = load_golden(case)
    if K:
        spec.sampler.K = K
    if T:
        spec.sampler.T = T
    return z, spec


def _U(z, spec):
    return torch.from_numpy(np.resize(z['in.U'], (spec.sampler.T, spec.sampler.nu)).copy())


@pytest.mark.parametrize('precision', PRECISIONS)
@pytest.mark.parametrize('case,K', [('push_nominal', 8192), ('peg_nominal', 32768), ('push_recovery', 4096),
                                    ('push_mlp_notransform', 5000), ('pendulum', 3000)])
def test_oracle_parity_at_benchmark_sizes(cuda_lib, case, K, precision):
    """Injected noise at the config's K (BASELINE.json configs 2-3): costs, argmin, action against the CPU oracle."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec(case, K)
    noise = O.sample_noise(spec, 99)
    U0 = _U(z, spec)
    want = O.command(spec, torch.from_numpy(z['in.state']), U0, noise)
    m = B200MPPI(spec, precision=precision, U_init=U0)
    a = m.command(z['in.state'], noise=noise.numpy())
    cost = m.cost_total
    rel = float(((cost - want['cost_total']).abs() / want['cost_total'].abs()).max())
    assert rel < RTOL[precision]
    # identical argmin is meaningful when the two best samples are further apart than the numeric error
    s = torch.sort(want['cost_total']).values
    gap = float(s[1] - s[0]) / float(s[0].abs())
    if gap > 10 * RTOL[precision]:
        assert int(cost.argmin()) == want['argmin'] == m.stats['argmin']
        assert float((a - want['action']).abs().max()) < 1e-4
    else:
        print('argmin / action assert skipped: top-2 cost gap', gap, case, K, precision)
    m.close()


@pytest.mark.parametrize('precision', ['f64', 'tf32x3', 'f16x3'])
def test_philox_path_costs_match_oracle_on_the_device_draw(cuda_lib, precision):
    """Production path: the device draws the noise; its own perturbed actions re-scored by the oracle."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', 4096)
    U0 = _U(z, spec)
    m = B200MPPI(spec, precision=precision, U_init=U0, seed=7)
    a = m.command(z['in.state'])
    pa = m.perturbed_action
    Ush = O.shift_nominal(spec, U0.clone())
    npost = pa - Ush
    oc = O.rollout_costs(spec, torch.from_numpy(z['in.state']), pa) + torch.sum(
        pa * (spec.sampler.lambda_ * npost @ torch.from_numpy(spec.sampler.noise_sigma_inv)), dim=(1, 2))
    cost = m.cost_total
    assert float(((cost - oc).abs() / oc.abs()).max()) < RTOL[precision]
    # softmax update recomputed on the host from those costs and that noise
    beta = oc.min()
    w = torch.exp(-(cost - cost.min()) / spec.sampler.lambda_)
    Unew = Ush + torch.einsum('k,kti->ti', w / w.sum(), npost)
    np.testing.assert_allclose(m.U.numpy(), Unew.numpy(), rtol=0, atol=1e-9)
    np.testing.assert_allclose(a.numpy(), Unew[0].numpy(), rtol=0, atol=1e-9)
    # bounds respected, every sample differs
    assert float(pa.min()) >= float(spec.sampler.u_min.min()) and float(pa.max()) <= float(spec.sampler.u_max.max())
    m.close()


def test_philox_noise_statistics(cuda_lib):
    """N(mu, Sigma) with the reference's covariance (push: mu=(0,0.1,0), Sigma=diag(0.2,0.4,0.7)); bounds lifted."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', 65536, T=8)
    spec.sampler.u_min = spec.sampler.u_max = None
    spec.sampler.noise_sigma = np.array([[0.2, 0.05, 0.0], [0.05, 0.4, -0.1], [0.0, -0.1, 0.7]])
    U0 = torch.zeros(8, 3, dtype=torch.double)
    m = B200MPPI(spec, precision='tf32x3', U_init=U0, seed=3)
    m.command(z['in.state'])
    pa = m.perturbed_action.reshape(-1, 3).numpy()   # U shifted = u_init rows only at the tail; U0 = 0
    pa = m.perturbed_action[:, :7].reshape(-1, 3).numpy()
    n = pa.shape[0]
    np.testing.assert_allclose(pa.mean(0), spec.sampler.noise_mu, atol=5 * np.sqrt(0.7 / n))
    np.testing.assert_allclose(np.cov(pa.T), spec.sampler.noise_sigma, atol=0.01)
    # independent across samples and steps, reproducible for the same (seed, call)
    m2 = B200MPPI(spec, precision='tf32x3', U_init=U0, seed=3)
    m2.command(z['in.state'])
    np.testing.assert_array_equal(m2.perturbed_action.numpy(), m.perturbed_action.numpy())
    m2.command(z['in.state'])
    assert np.abs(m2.perturbed_action.numpy() - m.perturbed_action.numpy()).max() > 0.1
    m.close(); m2.close()


@pytest.mark.parametrize('precision', ['f64', 'tf32x3', 'f16x3'])
def test_sample_shards_reproduce_the_unsharded_costs(cuda_lib, precision):
    """K-shard invariance on one GPU: two handles owning [0,K/2) and [K/2,K) see the same samples (global Philox
    index) and produce the same per-sample costs as one handle owning all K — at the 1M-sample size of config 5."""
    from tampc_b200.mppi import B200MPPI
    K = 1 << 17 if precision == 'f64' else 1 << 20
    z, spec = _spec('push_nominal', K, T=10)
    U0 = _U(z, spec)
    full = B200MPPI(spec, precision=precision, U_init=U0, seed=11)
    full.command(z['in.state'])
    c_full = full.cost_total.numpy()
    parts = []
    for r in range(2):
        h = B200MPPI(spec, precision=precision, U_init=U0, seed=11, rank=r, world_size=2)
        # single process: run only the local rollout (begin) and read the shard's costs
        import ctypes as C
        from tampc_b200 import _lib as L
        buf = torch.zeros(L.PARTIAL_DOUBLES, dtype=torch.double, device='cuda')
        x = np.ascontiguousarray(z['in.state'])
        h._check(h._lib.tampc_mppi_command_begin(h._h, L.dptr(x), L.NOISE_PHILOX, None, 0, C.c_void_p(buf.data_ptr())))
        h.synchronize()
        parts.append(h.cost_total.numpy())
        h.close()
    np.testing.assert_array_equal(np.concatenate(parts), c_full)
    assert np.all(np.isfinite(c_full)) and c_full.min() > 0
    full.close()


def test_commands_are_deterministic_and_advance_the_stream(cuda_lib):
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', 8192)
    runs = []
    for _ in range(2):
        m = B200MPPI(spec, precision='tf32x3', U_init=_U(z, spec), seed=5)
        acts = [m.command(z['in.state']).numpy() for _ in range(3)]
        runs.append((np.stack(acts), m.U.numpy()))
        m.close()
    np.testing.assert_array_equal(runs[0][0], runs[1][0])
    np.testing.assert_array_equal(runs[0][1], runs[1][1])
    assert np.abs(runs[0][0][0] - runs[0][0][1]).max() > 0


@pytest.mark.parametrize('precision', ['f64', 'tf32x3'])
def test_rollout_costs_and_get_rollouts_match_oracle(cuda_lib, precision):
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', 512)
    U0 = _U(z, spec)
    m = B200MPPI(spec, precision=precision, U_init=U0)
    g = torch.Generator().manual_seed(0)
    actions = torch.rand(512, 12, 3, dtype=torch.double, generator=g) * 2 - 1
    actions[:, :, 1] = actions[:, :, 1].abs()
    cost, states, acts = m._compute_rollout_costs(actions, state=z['in.state'])
    oc, ostates = O.rollout_costs(spec, torch.from_numpy(z['in.state']), actions, return_states=True)
    assert states.shape == (1, 512, 12, 5) and acts.shape == (1, 512, 12, 3)
    assert float(((cost - oc).abs() / oc.abs()).max()) < RTOL[precision]
    # float32-level modes: pose dims (what the cost sees) to 2e-5; the reaction-force dims swing by several units
    # per step through a 1/0.12 un-scaling and are only good to ~1e-4 of their magnitude
    tol_pose, tol_force = (1e-9, 1e-8) if precision == 'f64' else (2e-5, 1e-3)
    np.testing.assert_allclose(states[0, :, :, :3].numpy(), ostates[:, :, :3].numpy(), rtol=0, atol=tol_pose)
    np.testing.assert_allclose(states[0, :, :, 3:].numpy(), ostates[:, :, 3:].numpy(), rtol=0, atol=tol_force)
    roll = m.get_rollouts(z['in.state'])[0]
    want = O.get_rollouts(spec, z['in.state'], U0)
    np.testing.assert_allclose(roll[:, :3].numpy(), want[:, :3].numpy(), rtol=0, atol=tol_pose)
    np.testing.assert_allclose(roll[:, 3:].numpy(), want[:, 3:].numpy(), rtol=0, atol=tol_force)
    m.close()


def test_horizon_change_reset_and_cost_swap(cuda_lib):
    """What OnlineMPPI does around recovery mode: change_horizon(5), swap the cost program, change back
    (online_controller.py:400-402,416-418)."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', 1024)
    zr, spec_r = _spec('push_recovery', 1024)
    m = B200MPPI(spec, precision='f64', U_init=_U(z, spec))
    U_before = m.U.clone()
    m.change_horizon(5)
    assert m.T == 5 and torch.equal(m.U, U_before[:5])
    m.set_cost(spec_r.cost)
    noise = O.sample_noise(spec_r, 4, K=1024, T=5)
    spec_r.sampler.K = 1024
    a = m.command(zr['in.state'], noise=noise.numpy())
    spec_mix = S.RolloutSpec(spec.dynamics, spec_r.cost, spec_r.sampler)
    want = O.command(spec_mix, torch.from_numpy(zr['in.state']), U_before[:5], noise)
    assert float(((m.cost_total - want['cost_total']).abs() / want['cost_total'].abs()).max()) < 1e-9
    assert float((a - want['action']).abs().max()) < 1e-8
    m.change_horizon(40)
    assert m.T == 40 and m.U.shape == (40, 3)
    m.set_cost(spec.cost)
    m.reset()
    assert m.U.shape == (40, 3)
    m.command(z['in.state'])
    m.close()


def test_unsupported_shapes_and_bad_arguments_raise(cuda_lib):
    from tampc_b200.mppi import B200MPPI, TampcError
    z, spec = _spec('push_mlp_notransform', 256)
    rng = np.random.RandomState(0)
    spec.dynamics.net = S.MLPSpec([rng.randn(24, 8), rng.randn(24, 24), rng.randn(5, 24)],
                                  [np.zeros(24), np.zeros(24), np.zeros(5)])
    m = B200MPPI(spec, precision='tf32x3')  # widths without a tuned kernel: tcgen05 (padded to 16) here, the run-time-shaped kernel
    assert 'tcgen05' in m.kernel_name       # for other depths and precisions (tests/test_gpu_tc.py, tests/test_gpu_generic.py)
    m.close()
    m = B200MPPI(spec, precision='f64')
    assert 'generic' in m.kernel_name
    m.close()
    spec.dynamics.net = S.MLPSpec([rng.randn(300, 8), rng.randn(5, 300)], [np.zeros(300), np.zeros(5)])
    with pytest.raises(TampcError) as e:
        B200MPPI(spec, precision='tf32x3')
    assert e.value.code == -2 and 'wider than 256' in str(e.value)
    z, spec = _spec('push_nominal', 256)
    m = B200MPPI(spec, precision='tf32x3', U_init=_U(z, spec))
    with pytest.raises(ValueError):
        m.command(np.zeros(4))
    with pytest.raises(ValueError):
        m.command(z['in.state'], noise=np.zeros((3, 3, 3)))
    with pytest.raises(ValueError):
        m.U = torch.zeros(500, 3)
    m.close()


@pytest.mark.parametrize('precision', ['f64', 'tf32x3', 'mixed'])
@pytest.mark.parametrize('case', ['push_nominal', 'peg_nominal', 'pendulum', 'push_badstate'])
def test_predict_next_state_matches_oracle(cuda_lib, case, precision):
    """controller.py:312-313 / online_controller.py:47-58: one nominal-dynamics step for the chosen action."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec(case, 64)
    m = B200MPPI(spec, precision=precision, U_init=_U(z, spec))
    g = torch.Generator().manual_seed(2)
    for _ in range(3):
        u = torch.rand(spec.sampler.nu, dtype=torch.double, generator=g)
        x = torch.from_numpy(z['in.state']) if case == 'push_badstate' else torch.from_numpy(z['in.state']) + 0.05 * torch.randn(
            spec.dynamics.nx, dtype=torch.double, generator=g)
        want, _ = O.apply_dynamics(spec.dynamics, x.view(1, -1), u.view(1, -1))
        got = m.predict_next_state(x, u)
        assert got.shape == (1, spec.dynamics.nx)
        np.testing.assert_allclose(got, want.numpy(), rtol=0, atol=1e-10 if precision == 'f64' else 2e-4)
    m.close()


def test_noise_and_actions_attributes(cuda_lib):
    """`.noise` is the POST-clamp perturbation (controller.py:394) and `.actions` the (1,K,T,nu) perturbed actions."""
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden('push_nominal')
    m = B200MPPI(spec, precision='f64', U_init=torch.from_numpy(z['in.U']))
    m.command(z['in.state'], noise=z['in.noise'])
    want = O.command(spec, torch.from_numpy(z['in.state']), torch.from_numpy(z['in.U']), torch.from_numpy(z['in.noise']))
    np.testing.assert_allclose(m.noise.numpy(), want['noise'].numpy(), rtol=0, atol=1e-12)
    np.testing.assert_allclose(m.actions.numpy()[0], want['perturbed_action'].numpy(), rtol=0, atol=1e-12)
    # clamped entries exist in this fixture: the post-clamp noise differs from the injected draw there
    assert np.any(np.abs(m.noise.numpy() - z['in.noise']) > 1e-6)
    m.close()


@pytest.mark.parametrize('precision', ['f64', 'tf32x3', 'f16x3', 'mixed'])
@pytest.mark.parametrize('K,T,P', [(1, 40, 3), (17, 7, 0), (33, 128, 16), (4737, 3, 64), (12289, 5, 1)])
def test_ragged_and_maximum_sizes(cuda_lib, precision, K, T, P):
    """Edge sizes: a single sample, K that is not a multiple of any tile (idle rows / warps / tiles in every kernel
    family and on both sides of the role-split / single-warp switch-overs), the longest horizon (TAMPC_MAX_HORIZON) and
    the largest trap set (TAMPC_MAX_TRAPS), an empty trap set."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', K, T)
    rng = np.random.RandomState(K)
    spec.cost.trap_x = np.tile(spec.cost.trap_x[:1], (P, 1)) + 0.3 * rng.randn(P, 5) if P else np.zeros((0, 5))
    spec.cost.trap_u = np.tile(spec.cost.trap_u[:1], (P, 1)) + 0.3 * rng.randn(P, 3) if P else np.zeros((0, 3))
    U0 = _U(z, spec)
    noise = O.sample_noise(spec, 17, K=K, T=T)
    want = O.command(spec, torch.from_numpy(z['in.state']), U0, noise)
    m = B200MPPI(spec, precision=precision, U_init=U0, max_horizon=128)
    a = m.command(z['in.state'], noise=noise.numpy())
    cost = m.cost_total
    assert cost.shape == (K,)
    rel = float(((cost - want['cost_total']).abs() / want['cost_total'].abs()).max())
    # the 1e-5 gate is for the task horizons; over 128 steps the float32-class modes drift further (measured on B200: tf32x3
    # 2.8e-5, mixed 1.2e-5, f16x3 8.9e-6; f64 5e-14) — DESIGN.md §4 records it as a known limit
    assert rel < (RTOL[precision] if T <= 40 or precision == 'f64' else 3e-5), rel
    s = torch.sort(want['cost_total']).values
    gap = 1.0 if K == 1 else float(s[1] - s[0]) / float(s[0].abs())
    print('K', K, 'T', T, 'P', P, precision, 'cost rel', rel, 'top-2 cost gap', gap)
    if gap > 1e-3:
        assert m.stats['argmin'] == want['argmin']
        assert float((a - want['action']).abs().max()) < 1e-4
    else:
        print('argmin / action assert skipped: top-2 cost gap', gap)
    m.close()


@pytest.mark.parametrize('precision', ['tf32x3', 'f16x3'])
@pytest.mark.parametrize('K', [56833, 65536, 66304, 66305])
def test_balanced_wave_kernels(cuda_lib, monkeypatch, precision, K):
    """56 832 < K <= 66 304 runs the one-CTA-per-SM kernel with seven tiles per sub-partition (one- and two-tile warps
    in one CTA; sample order sub-partition-major, the last CTA ragged), above it the 16-warp wave.  Oracle parity, and
    agreement with the default kernels to float32 rounding (a one-tile warp sums a sample's running cost over two
    lanes, a two-tile warp over one: the order of the float32 additions differs, nothing else)."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', K, 5)
    U0 = _U(z, spec)
    noise = O.sample_noise(spec, 23, K=K, T=5)
    want = O.command(spec, torch.from_numpy(z['in.state']), U0, noise)
    m = B200MPPI(spec, precision=precision, U_init=U0)
    a = m.command(z['in.state'], noise=noise.numpy())
    cost = m.cost_total.clone()
    assert float(((cost - want['cost_total']).abs() / want['cost_total'].abs()).max()) < RTOL[precision]
    assert m.stats['argmin'] == want['argmin']
    assert float((a - want['action']).abs().max()) < 1e-4
    monkeypatch.setenv('TAMPC_SM_WAVE', '0')
    m.U = U0.clone()
    m.command(z['in.state'], noise=noise.numpy())
    assert float(((m.cost_total - cost).abs() / cost.abs()).max()) < 1e-6
    m.close()


@pytest.mark.parametrize('precision', ['tf32x3', 'f64'])
def test_opt_in_fused_update_matches_the_default_tail(cuda_lib, monkeypatch, precision):
    """TAMPC_FUSE=1 (with the role-split kernels off) runs the softmax update inside the rollout launch and publishes
    the result block from there; action, nominal sequence and stats must agree with the two-launch default."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', 500)
    U0 = _U(z, spec)
    noise = O.sample_noise(spec, 5, K=500)

    def run():
        m = B200MPPI(spec, precision=precision, U_init=U0)
        a = m.command(z['in.state'], noise=noise.numpy())
        out = (a.clone(), m.U.clone(), dict(m.stats), m.stats['launches'])
        m.close()
        return out

    a0, U_0, st0, n0 = run()
    monkeypatch.setenv('TAMPC_SPLIT', '0')
    monkeypatch.setenv('TAMPC_FUSE', '1')
    a1, U_1, st1, n1 = run()
    assert n1 < n0  # one launch instead of two: the fused path really ran
    assert st1['argmin'] == st0['argmin']
    assert float((a1 - a0).abs().max()) < 1e-6
    assert float((U_1 - U_0).abs().max()) < 1e-6


@pytest.mark.parametrize('precision', ['tf32x3', 'f16x3'])
def test_persistent_multiwave_kernel_same_bits(cuda_lib, monkeypatch, precision):
    """K beyond one wave: the slot-major persistent launch (rollout_mma_persist_kernel) and the plain grid launch run the
    same per-sample arithmetic — bitwise-identical costs, ragged last unit included."""
    from tampc_b200.mppi import B200MPPI
    z, spec = _spec('push_nominal', 32 * 148 * 16 * 2 + 32 * 148 * 5 + 21, T=4)  # 2 passes + a partial third, ragged last unit
    costs = []
    for persist in ('0', '1'):
        monkeypatch.setenv('TAMPC_PERSIST', persist)
        m = B200MPPI(spec, precision=precision, U_init=_U(z, spec), seed=3)
        assert ('persist' in m.kernel_name) == (persist == '1'), m.kernel_name
        m.command(z['in.state'])
        costs.append(m.cost_total.clone())
        m.close()
    assert torch.equal(costs[0], costs[1])
