"""GPU: the device hyper-parameter fit of the online GP (tampc_gp_fit) against the trajectory of the
reference's OnlineGPMixing._fit_params — Adam (lr 0.1) through the restated gpytorch — recorded in
tests/golden/gp_fit.npz by oracle/make_golden.py: 150 iterations on 30 points at construction, then four updates of 15
iterations each with the optimiser state carried over."""
import ctypes as C
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
import os

pytestmark = pytest.mark.gpu

# Adam divides by sqrt(v): 150 + 60 float64 iterations of two different gradient evaluations (autograd through a
# Cholesky vs the analytic trace form) stay this close
RAW_ATOL = 2e-6


def _golden():
    return np.load(os.path.join(GOLDEN_DIR, 'gp_fit.npz'))


def test_fit_trajectory_matches_reference(cuda_lib):
    from tampc_b200 import _lib as L
    z = _golden()
    ny = z['stage0.raw_lengthscale'].shape[0]
    raw = np.zeros(3 * ny + 1)
    m, v, step = np.zeros(3 * ny + 1), np.zeros(3 * ny + 1), C.c_int64(0)
    loss = np.zeros(2)
    total = 0
    for s in range(int(z['n_stages'])):
        X = np.ascontiguousarray(z['stage%d.X' % s])
        R = np.ascontiguousarray(z['stage%d.resid' % s])
        iters = int(z['stage%d.iters' % s])
        rc = cuda_lib.tampc_gp_fit(0, X.shape[0], X.shape[1], ny, L.dptr(X), L.dptr(R), iters, 0.1, L.dptr(raw), L.dptr(m), L.dptr(v),
                                   C.byref(step), L.dptr(loss))
        assert rc == 0, cuda_lib.tampc_gp_fit_last_error()
        total += iters
        assert step.value == total
        want = np.concatenate([z['stage%d.raw_lengthscale' % s], z['stage%d.raw_outputscale' % s], z['stage%d.raw_task_noises' % s],
                               z['stage%d.raw_noise' % s]])
        np.testing.assert_allclose(raw, want, rtol=0, atol=RAW_ATOL, err_msg='stage %d' % s)
        np.testing.assert_allclose(loss[0], float(z['stage%d.last_loss' % s]), rtol=1e-7, atol=1e-7)


def test_fit_rejects_bad_arguments(cuda_lib):
    from tampc_b200 import _lib as L
    X, R = np.zeros((60, 8)), np.zeros((60, 5))
    raw, m, v, step, loss = np.zeros(16), np.zeros(16), np.zeros(16), C.c_int64(0), np.zeros(2)
    rc = cuda_lib.tampc_gp_fit(0, 60, 8, 5, L.dptr(X), L.dptr(R), 1, 0.1, L.dptr(raw), L.dptr(m), L.dptr(v), C.byref(step), L.dptr(loss))
    assert rc == -2 and b'TAMPC_GP_FIT_MAX_POINTS' in cuda_lib.tampc_gp_fit_last_error()


def test_device_fitted_online_gp_feeds_the_rollout(cuda_lib):
    """B200OnlineGP in place of OnlineGPMixing: constructor fit, `update` bookkeeping + refit, and the fitted object read
    by extract.py like the original."""
    import fake_tampc
    from tampc_b200 import extract, spec as S
    from tampc_b200.online_gp import make_online_gp_class
    z = _golden()
    ny = 5
    X0, R0 = torch.from_numpy(z['stage0.X']), torch.from_numpy(z['stage0.resid'])
    pre = fake_tampc.identity_preprocessor()
    ds = types.SimpleNamespace(training_set=lambda: (X0, R0, None), config=types.SimpleNamespace(ny=ny, nx=5, nu=3),
                               original_config=lambda: types.SimpleNamespace(predict_difference=True, predict_all_dims=True, nx=5),
                               preprocessor=pre)
    prior = types.SimpleNamespace(get_batch_predictions=lambda xu, all_in_transformed_space=True: torch.zeros(len(xu), ny, dtype=torch.double),
                                  dyn_net=None)
    cls = make_online_gp_class(fake_tampc.fake_online_model_module())
    om = cls(prior, ds, lambda a, b: a - b, slice_to_use=slice(0, 30), training_iter=150)
    assert om.gp.training is False and len(om.xu) == 30
    np.testing.assert_allclose(om.raw_lengthscale, z['stage0.raw_lengthscale'], atol=RAW_ATOL)
    np.testing.assert_allclose(om.raw_noise, z['stage0.raw_noise'], atol=RAW_ATOL)
    assert abs(om.last_loss - float(z['stage0.last_loss'])) < 1e-6
    for s in range(1, int(z['n_stages'])):
        xu_new, y_new = torch.from_numpy(z['stage%d.X' % s][-1]), torch.from_numpy(z['stage%d.resid' % s][-1])
        om._update(xu_new[:5], xu_new[5:], y_new)
        assert len(om.xu) == 30 + s
        np.testing.assert_allclose(om.raw_outputscale, z['stage%d.raw_outputscale' % s], atol=RAW_ATOL)
        np.testing.assert_allclose(om.raw_task_noises, z['stage%d.raw_task_noises' % s], atol=RAW_ATOL)
    g = extract.gp_from_online_model(om)
    sp = np.log1p(np.exp(z['stage4.raw_lengthscale']))
    np.testing.assert_allclose(g.lengthscale, sp, rtol=1e-5)
    np.testing.assert_allclose(g.noise, np.log1p(np.exp(z['stage4.raw_task_noises'])) + np.log1p(np.exp(z['stage4.raw_noise'])) + 2e-4, rtol=1e-5)
    assert g.X.shape == (34, 8) and isinstance(g, S.GPSpec)
    with pytest.raises(RuntimeError):
        om.predict(None, None, X0[:1, :5], X0[:1, 5:])
