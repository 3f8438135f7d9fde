"""GPU parity of the online GP residual: the (sample, replica) kernel behind the C ABI against the
reference's OnlineGPMixing outputs (tests/golden/push_gp_*.npz) with the action noise AND the GP draws injected."""
import numpy as np
import pytest
import torch

from conftest import GP_CASES, load_golden

pytestmark = pytest.mark.gpu

# float64 runs the reference's arithmetic.  The float32 modes keep the NETWORKS in float32 but evaluate the GP posterior
# (kernel vector, mean, variance, draw) in float64 (gp_rollout.cuh), so they are held to the north-star gate of the
# deterministic path: per-sample cost 1e-5 relative.
COST_RTOL = {'f64': 1e-9, 'tf32x3': 1e-5, 'mixed': 1e-5, 'f32': 1e-5}
ACTION_ATOL = {'f64': 1e-8, 'tf32x3': 1e-4, 'mixed': 1e-4, 'f32': 1e-4}


def _rel(a, b):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize('precision', ['f64', 'tf32x3', 'mixed', 'f32'])
@pytest.mark.parametrize('case', GP_CASES)
def test_gp_command_matches_reference_golden(cuda_lib, case, precision):
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden(case)
    mppi = B200MPPI(spec, precision=precision, U_init=torch.from_numpy(z['in.U']))
    action = mppi.command(z['in.state'], noise=z['in.noise'], gp_eps=z['in.gp_eps']).numpy()
    cost = mppi.cost_total.numpy()
    assert np.all(np.isfinite(cost))
    rel = _rel(cost, z['ref.cost_total'])
    print(case, precision, 'cost rel', rel, 'action err', np.max(np.abs(action - z['ref.action'])))
    assert rel < COST_RTOL[precision]
    assert int(np.argmin(cost)) == int(z['ref.argmin'])
    assert np.max(np.abs(action - z['ref.action'])) < ACTION_ATOL[precision]
    assert np.max(np.abs(mppi.U.numpy() - z['ref.U'])) < ACTION_ATOL[precision]
    mppi.close()


@pytest.mark.parametrize('case', GP_CASES)
def test_gp_get_rollouts_at_the_mean(cuda_lib, case):
    """With zero draws the nominal sequence rolls out along the posterior mean (ref.get_rollouts was made that way)."""
    from tampc_b200.mppi import B200MPPI
    from tampc_b200 import _lib as L
    z, spec = load_golden(case)
    mppi = B200MPPI(spec, precision='f64', U_init=torch.from_numpy(z['ref.U']))
    ny = spec.dynamics.gp.resid.shape[1]
    zeros = np.zeros((mppi.K_local, mppi.M, mppi.T, ny))
    mppi._check(mppi._lib.tampc_mppi_set_gp_draws(mppi._h, L.dptr(zeros), zeros.size))
    roll = mppi.get_rollouts(z['in.state'])[0].numpy()
    np.testing.assert_allclose(roll, z['ref.get_rollouts'], rtol=1e-9, atol=1e-9)
    mppi.close()


def test_gp_mean_only_is_replica_invariant(cuda_lib):
    """sample_dynamics=False makes the model deterministic again: M=10 and M=1 give the same costs."""
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden('push_gp_local')
    spec.dynamics.gp.sample = False
    out = []
    for M in (10, 1, 7):
        spec.sampler.rollout_samples = M
        m = B200MPPI(spec, precision='f64', U_init=torch.from_numpy(z['in.U']))
        m.command(z['in.state'], noise=z['in.noise'])
        out.append(m.cost_total.numpy())
        m.close()
    np.testing.assert_allclose(out[0], out[1], rtol=1e-12)
    np.testing.assert_allclose(out[0], out[2], rtol=1e-12)


def test_gp_matches_oracle_on_fresh_draws(cuda_lib):
    """Same fixtures, new seeds, larger K: kernel vs CPU oracle with both noise sources injected."""
    from oracle import mppi_oracle as O
    from tampc_b200.mppi import B200MPPI
    for case in GP_CASES:
        z, spec = load_golden(case)
        spec.sampler.K = 96
        g = torch.Generator().manual_seed(7)
        noise = O.sample_noise(spec, 99)
        ny = spec.dynamics.gp.resid.shape[1]
        eps = torch.randn(spec.sampler.T, spec.sampler.rollout_samples, spec.sampler.K, ny, dtype=torch.double, generator=g)
        want = O.command(spec, torch.from_numpy(z['in.state']), torch.from_numpy(z['in.U']), noise, eps)
        m = B200MPPI(spec, precision='f64', U_init=torch.from_numpy(z['in.U']))
        action = m.command(z['in.state'], noise=noise, gp_eps=eps).numpy()
        assert _rel(m.cost_total.numpy(), want['cost_total'].numpy()) < 1e-9
        assert m.stats['argmin'] == want['argmin']
        np.testing.assert_allclose(action, want['action'].numpy(), atol=1e-8)
        m.close()


def test_gp_philox_draws_are_deterministic_and_spread(cuda_lib):
    """On-device draws: keyed by (seed, call, global sample, replica, t) — same handle state gives the same costs; the
    sampled costs scatter around the mean-only costs."""
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden('push_gp_latent')
    costs = []
    for _ in range(2):
        m = B200MPPI(spec, precision='mixed', U_init=torch.from_numpy(z['in.U']), seed=5)
        m.command(z['in.state'], noise=z['in.noise'])
        costs.append(m.cost_total.numpy())
        m.close()
    np.testing.assert_array_equal(costs[0], costs[1])
    spec.dynamics.gp.sample = False
    m = B200MPPI(spec, precision='mixed', U_init=torch.from_numpy(z['in.U']), seed=5)
    m.command(z['in.state'], noise=z['in.noise'])
    mean_cost = m.cost_total.numpy()
    m.close()
    assert np.any(costs[0] != mean_cost)
    assert np.median(np.abs(costs[0] - mean_cost) / mean_cost) < 0.2


def test_gp_toggle_returns_to_the_tensor_core_path(cuda_lib):
    """set_gp(None) re-lays the networks out for the mma kernels: results equal a handle that never had a GP."""
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden('push_gp_local')
    gp = spec.dynamics.gp
    m = B200MPPI(spec, precision='tf32x3', U_init=torch.from_numpy(z['in.U']))
    m.set_gp(None)
    a1 = m.command(z['in.state'], noise=z['in.noise']).numpy()
    c1 = m.cost_total.numpy()
    m.close()
    spec.dynamics.gp = None
    m = B200MPPI(spec, precision='tf32x3', U_init=torch.from_numpy(z['in.U']))
    a2 = m.command(z['in.state'], noise=z['in.noise']).numpy()
    np.testing.assert_array_equal(c1, m.cost_total.numpy())
    np.testing.assert_array_equal(a1, a2)
    m.set_gp(gp)  # and back on
    m.U = torch.from_numpy(z['in.U'])
    m.command(z['in.state'], noise=z['in.noise'], gp_eps=z['in.gp_eps'])
    assert _rel(m.cost_total.numpy(), z["ref.cost_total"]) < 1e-5
    m.close()
