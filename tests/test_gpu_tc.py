"""GPU parity of the tcgen05 kernel (tampc_b200/csrc/tc_rollout.cuh): plain-MLP dynamics of ANY hidden widths (multiples of
16 up to 256) — what arm_pytorch_utilities.model.make.make_sequential_network(h_units=...) can build (tampc/util.py:560) —
on tcgen05.mma with TMEM operands and the TMA-fed weight ring, against the CPU oracle at the north-star gates: per-sample
cost 1e-5 relative, action 1e-4 absolute, identical argmin."""
import numpy as np
import pytest
import torch

from oracle import mppi_oracle as O
from tampc_b200.synthetic import make_synthetic_spec

pytestmark = pytest.mark.gpu


def _check(spec, x0, K, T, seed, monkeypatch=None):
    from tampc_b200.mppi import B200MPPI
    noise = O.sample_noise(spec, seed, K=K, T=T)
    U0 = torch.zeros(T, spec.sampler.nu, dtype=torch.double)
    want = O.command(spec, torch.from_numpy(x0), U0, noise)
    m = B200MPPI(spec, precision='tf32x3', U_init=U0, max_horizon=max(T, 64))
    assert 'tcgen05' in m.kernel_name, m.kernel_name
    a = m.command(x0, noise=noise.numpy())
    cost = m.cost_total
    assert torch.isfinite(cost).all()
    rel = float(((cost - want['cost_total']).abs() / want['cost_total'].abs()).max())
    s = torch.sort(want['cost_total']).values
    gap = float((s[1] - s[0]) / s[0].abs()) if K > 1 else 1.0
    print(spec.name, 'K', K, 'T', T, 'cost rel', rel, 'top-2 gap', gap)
    assert rel < 1e-5, rel
    if gap > 1e-4:
        assert int(cost.argmin()) == want['argmin'] == m.stats['argmin']
        assert float((a - want['action']).abs().max()) < 1e-4
    else:
        print('   argmin / action assert skipped: top-2 gap', gap)
    return m, want


@pytest.mark.parametrize('H1,H2', [(16, 16), (32, 32), (48, 32), (64, 96), (96, 64), (128, 128), (32, 256), (256, 64), (240, 176), (256, 256)])
def test_any_hidden_widths_match_the_oracle(cuda_lib, monkeypatch, H1, H2):
    """Widths without a tuned kernel go to the tcgen05 kernel by themselves; the tuned shapes (32, 64, 128, 256) are
    forced onto it with TAMPC_TC=1.  Ragged K (idle rows in the last 128-sample tile)."""
    monkeypatch.setenv('TAMPC_TC', '1')
    K, T = 1000, 12
    spec, x0 = make_synthetic_spec(H=H1, H2=H2, K=K, T=T, seed=H1 + H2)
    m, want = _check(spec, x0, K, T, 5)
    # get_rollouts / predict_next_state: the same kernel with one live sample
    roll = m.get_rollouts(x0)[0]
    ref = O.get_rollouts(spec, torch.from_numpy(x0), m.U)
    assert float((roll - ref).abs().max()) < 1e-4
    # the production path: on-device Philox noise, the device's own perturbed actions re-scored by the oracle
    m.U = torch.zeros(T, 3, dtype=torch.double)
    m.command(x0)
    pa = m.perturbed_action
    oc = O.rollout_costs(spec, torch.from_numpy(x0), pa)
    Ush = O.shift_nominal(spec, torch.zeros(T, 3, dtype=torch.double))
    oc = oc + torch.sum(pa * (spec.sampler.lambda_ * (pa - Ush) @ torch.from_numpy(spec.sampler.noise_sigma_inv)), dim=(1, 2))
    assert float(((m.cost_total - oc).abs() / oc.abs()).max()) < 1e-5
    m.close()


@pytest.mark.parametrize('H,K,T', [(256, 4096, 40), (128, 2049, 40), (256, 130, 100), (64, 1, 5), (256, 128 * 148 + 7, 3)])
def test_wide_layers_at_task_horizons_and_odd_sizes(cuda_lib, monkeypatch, H, K, T):
    """H = 256 streams W2 (512 KB of hi/lo pieces) through the TMA ring every step; H = 128 keeps it resident.  A single
    sample, a horizon of 100, and more tiles than SMs."""
    monkeypatch.setenv('TAMPC_TC', '1')
    spec, x0 = make_synthetic_spec(H=H, K=K, T=T, seed=7 * H + T)
    m, _ = _check(spec, x0, K, T, 9)
    m.close()


def test_default_choice_keeps_the_tuned_kernels_for_their_shapes(cuda_lib):
    from tampc_b200.mppi import B200MPPI
    spec, _ = make_synthetic_spec(H=32, K=512, T=10)
    m = B200MPPI(spec, precision='tf32x3')
    assert 'tcgen05' not in m.kernel_name
    m.close()
    spec, _ = make_synthetic_spec(H=48, K=512, T=10)
    m = B200MPPI(spec, precision='tf32x3')
    assert 'tcgen05' in m.kernel_name
    m.close()


def test_widths_outside_the_kernel_go_elsewhere_or_are_refused(cuda_lib):
    from tampc_b200.mppi import B200MPPI, TampcError
    for H1, H2 in [(40, 32), (32, 8), (24, 100)]:   # not multiples of 16: padded with zero units, same kernel, same gates
        spec, x0 = make_synthetic_spec(H=H1, H2=H2, K=700, T=20, seed=H1 + H2)
        m, _ = _check(spec, x0, 700, 20, 3)
        assert 'tcgen05' in m.kernel_name
        m.close()
    spec, _ = make_synthetic_spec(H=272, H2=32, K=256, T=10)
    with pytest.raises(TampcError) as e:
        B200MPPI(spec, precision='tf32x3')
    assert e.value.code == -2
    spec, _ = make_synthetic_spec(H=48, K=256, T=10)
    m = B200MPPI(spec, precision='f64')  # the tcgen05 kernel is 3xTF32 only: float64 of the same widths is the generic kernel
    assert 'generic' in m.kernel_name and 'f64' in m.kernel_name
    m.close()

