"""CPU: the oracle's GP-residual restatement against golden vectors produced by the reference's
own OnlineGPMixing / OnlineDynamicsModel.predict run over oracle/shim (gpytorch itself is not available: its exact-GP
arithmetic is restated in oracle/shim/gpytorch, parity unpinned upstream)."""
import numpy as np
import pytest
import torch

from conftest import GP_CASES, load_golden
from oracle import mppi_oracle as O
from tampc_b200 import spec as S


@pytest.mark.parametrize('case', GP_CASES)
def test_oracle_reproduces_reference_gp_command(case):
    z, spec = load_golden(case)
    assert spec.dynamics.gp is not None and int(z['ref.M']) == spec.sampler.rollout_samples == 10
    out = O.command(spec, torch.from_numpy(z['in.state']), torch.from_numpy(z['in.U']), torch.from_numpy(z['in.noise']),
                    torch.from_numpy(z['in.gp_eps']))
    ref_cost = torch.from_numpy(z['ref.cost_total'])
    assert ((out['cost_total'] - ref_cost).abs() / ref_cost.abs()).max().item() < 1e-11
    assert out['argmin'] == int(z['ref.argmin'])
    np.testing.assert_allclose(out['action'].numpy(), z['ref.action'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out['U'].numpy(), z['ref.U'], rtol=0, atol=1e-12)


@pytest.mark.parametrize('case', GP_CASES)
def test_oracle_gp_states_per_replica(case):
    """ref.states_first8 = states[replica 0, first 8 samples] of the reference's (M,K,T,nx) stack."""
    z, spec = load_golden(case)
    U = O.shift_nominal(spec, torch.from_numpy(z['in.U']).clone())
    perturbed, _, _ = O.perturb(spec, U, torch.from_numpy(z['in.noise']))
    _, states = O.rollout_costs(spec, torch.from_numpy(z['in.state']), perturbed, return_states=True,
                                gp_eps=torch.from_numpy(z['in.gp_eps']))
    assert states.shape[:3] == (10, spec.sampler.K, spec.sampler.T)
    np.testing.assert_allclose(states[0, :8].numpy(), z['ref.states_first8'], rtol=1e-10, atol=1e-10)


def test_gp_spaces_of_the_two_fixtures():
    assert load_golden('push_gp_latent')[1].dynamics.gp.space == S.GP_AT_NETWORK   # nominal_adapt=GP_KERNEL_INDEP_OUT
    local = load_golden('push_gp_local')[1]
    assert local.dynamics.gp.space == S.GP_ORIGINAL and len(local.dynamics.gp.X) == 6  # temporary local model
    assert local.sampler.rollout_var_cost == 0.5


def test_gp_posterior_interpolates_its_data():
    """At a training input the posterior mean moves towards the target by K (K + nI)^-1 and the variance drops to
    about the noise level; far away it returns to the prior (mean function, s + n)."""
    _, spec = load_golden('push_gp_local')
    d, g = spec.dynamics, spec.dynamics.gp
    mean, var = O.gp_posterior(d, torch.from_numpy(g.X))
    assert np.all(var.numpy() < (g.outputscale + g.noise) * (1 + 1e-12))
    assert np.all(var.numpy() >= g.noise * (1 - 1e-9))
    far = torch.from_numpy(g.X[:1] + 1e3)
    mean_far, var_far = O.gp_posterior(d, far)
    np.testing.assert_allclose(mean_far.numpy(), 0, atol=1e-12)
    np.testing.assert_allclose(var_far.numpy()[0], g.outputscale + g.noise, rtol=1e-12)


def test_gp_spec_roundtrip():
    for case in GP_CASES:
        _, spec = load_golden(case)
        a = S.spec_to_arrays(spec)
        b = S.spec_to_arrays(S.spec_from_arrays(a))
        assert a.keys() == b.keys()
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])


def test_extraction_of_a_duck_typed_online_gp():
    """extract.py reads an OnlineGPMixing-like object by attribute names only (runs on the GPU box too)."""
    import fake_tampc
    from tampc_b200 import extract
    for case in GP_CASES:
        z, spec = load_golden(case)
        ctrl = fake_tampc.make_controller(spec, z['in.U'])
        nominal = ctrl.dynamics.nominal_model
        ctrl.dynamics.nominal_model = fake_tampc.make_online_gp(spec, nominal)
        got = extract.spec_from_controller(ctrl)
        a, b = S.spec_to_arrays(spec), S.spec_to_arrays(got)
        for k in a:
            if k.endswith('name') or 'bad_state_norm' in k:
                continue
            np.testing.assert_allclose(np.asarray(a[k], dtype=float), np.asarray(b[k], dtype=float), rtol=0, atol=1e-12, err_msg=k)
        ctrl.dynamics.nominal_model.gp.training = True
        with pytest.raises(extract.UnsupportedSpec):
            extract.spec_from_controller(ctrl)
        ctrl.dynamics.nominal_model.gp.training = False
        ctrl.dynamics.nominal_model.use_independent_outputs = False
        with pytest.raises(extract.UnsupportedSpec):
            extract.spec_from_controller(ctrl)
