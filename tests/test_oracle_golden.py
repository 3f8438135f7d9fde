"""CPU: the oracle restatement against the committed golden vectors (generated from the reference's own in-tree
classes by oracle/make_golden.py).  Runs anywhere; no reference tree and no GPU needed."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_CASES, load_golden
from oracle import mppi_oracle as O
from tampc_b200 import spec as S


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_oracle_reproduces_reference_outputs(case):
    z, spec = load_golden(case)
    out = O.command(spec, torch.from_numpy(z['in.state']), torch.from_numpy(z['in.U']), torch.from_numpy(z['in.noise']))
    ref_cost = torch.from_numpy(z['ref.cost_total'])
    rel = ((out['cost_total'] - ref_cost).abs() / ref_cost.abs()).max().item()
    assert rel < 1e-12
    assert out['argmin'] == int(z['ref.argmin'])
    np.testing.assert_allclose(out['action'].numpy(), z['ref.action'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out['U'].numpy(), z['ref.U'], rtol=0, atol=1e-12)
    np.testing.assert_allclose(out['omega'].numpy(), z['ref.omega'], rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_oracle_rollout_states_and_get_rollouts(case):
    z, spec = load_golden(case)
    U = O.shift_nominal(spec, torch.from_numpy(z['in.U']).clone())
    perturbed, _, _ = O.perturb(spec, U, torch.from_numpy(z['in.noise']))
    _, states = O.rollout_costs(spec, torch.from_numpy(z['in.state']), perturbed, return_states=True)
    np.testing.assert_allclose(states[:8].numpy(), z['ref.states_first8'], rtol=1e-12, atol=1e-12)
    roll = O.get_rollouts(spec, torch.from_numpy(z['in.state']), torch.from_numpy(z['ref.U']))
    np.testing.assert_allclose(roll.numpy(), z['ref.get_rollouts'], rtol=1e-12, atol=1e-12)


def test_injected_noise_convention_is_reproducible():
    z, spec = load_golden('push_nominal')
    noise = O.sample_noise(spec, int(z['in.noise_seed']))
    np.testing.assert_array_equal(noise.numpy(), z['in.noise'])


def test_replica_count_does_not_change_deterministic_costs():
    """push_M10 was produced by the reference at M=10; the oracle rolls one replica."""
    z, spec = load_golden('push_M10')
    assert int(z['ref.M']) == 10 and spec.sampler.rollout_samples == 10
    out = O.command(spec, torch.from_numpy(z['in.state']), torch.from_numpy(z['in.U']), torch.from_numpy(z['in.noise']))
    assert ((out['cost_total'].numpy() - z['ref.cost_total']) / z['ref.cost_total']).max() < 1e-12


def test_bad_state_guard_zeroes_state_and_action():
    """online_controller.py:62-66,77: rows with ||x|| > 1e5 come back as zeros and cost sees u = 0."""
    z, spec = load_golden('push_badstate')
    x = torch.from_numpy(z['in.state']).view(1, -1).repeat(3, 1)
    x[1] = 0.1
    u = torch.ones(3, spec.dynamics.nu, dtype=torch.double) * 0.3
    nxt, u_seen = O.apply_dynamics(spec.dynamics, x, u)
    assert torch.all(nxt[0] == 0) and torch.all(nxt[2] == 0) and torch.all(u_seen[0] == 0)
    assert torch.any(nxt[1] != 0) and torch.all(u_seen[1] == 0.3)


def test_spec_roundtrip_through_arrays():
    for case in GOLDEN_CASES:
        z, spec = load_golden(case)
        again = S.spec_from_arrays(S.spec_to_arrays(spec))
        a, b = S.spec_to_arrays(spec), S.spec_to_arrays(again)
        assert a.keys() == b.keys()
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])


def test_macs_per_step_match_survey():
    assert load_golden('push_nominal')[1].dynamics.macs_per_step == 3523
    assert load_golden('peg_nominal')[1].dynamics.macs_per_step == 3507


def test_packaged_spec_roundtrip(tmp_path):
    """a RolloutSpec saved beside the reference checkpoint reloads to the same problem."""
    for case in ('push_nominal', 'push_recovery', 'pendulum', 'push_gp_local'):
        z, spec = load_golden(case)
        path = str(tmp_path / (case + '.spec.npz'))
        S.save_spec(path, spec)
        again = S.load_spec(path)
        a, b = S.spec_to_arrays(spec), S.spec_to_arrays(again)
        assert a.keys() == b.keys()
        for k in a:
            np.testing.assert_array_equal(a[k], b[k])
    with pytest.raises(ValueError):
        S.load_spec(str(__import__('os').path.join(__import__('conftest').GOLDEN_DIR, 'push_nominal.npz')))  # not a package
