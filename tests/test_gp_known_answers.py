"""CPU: the exact-GP arithmetic that the a10 / f-2 goldens rest on, pinned INDEPENDENTLY of the repo's own restatements.

`oracle/shim/gpytorch` (which the unmodified reference `online_model.py` runs over, gpytorch itself being absent) and
`oracle.mppi_oracle.gp_posterior` are checked against a plain numpy / scipy implementation written here straight from
Rasmussen & Williams, "Gaussian Processes for Machine Learning" (2006):
    eq. 2.25  mean*  = k*^T (K + s_n^2 I)^-1 y
    eq. 2.26  var*   = k(x*,x*) - k*^T (K + s_n^2 I)^-1 k*
    eq. 2.30  log p(y|X) = -1/2 y^T (K + s_n^2 I)^-1 y - 1/2 log|K + s_n^2 I| - n/2 log 2 pi
    eq. 5.9   d/dtheta log p = 1/2 tr((alpha alpha^T - K^-1) dK/dtheta)
with scipy.linalg.cho_factor / cho_solve, and against finite differences of 2.30 for the gradient that Adam consumes
(the device fit tampc_gp_fit implements the trace form 5.9; tests/test_gpu_gp_fit.py compares it with the golden
trajectory produced through the shim's autograd)."""
import os
import sys

import numpy as np
import pytest
import scipy.linalg
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim'))
import gpytorch  # noqa: E402  (the shim)

from oracle import mppi_oracle as O  # noqa: E402
from tampc_b200 import spec as S  # noqa: E402


def softplus(x):
    return np.log1p(np.exp(x))


def rbf(A, B, ls, os_):
    d2 = ((A[:, None, :] - B[None, :, :]) ** 2).sum(-1)
    return os_ * np.exp(-0.5 * d2 / ls ** 2)


def rw_posterior(X, y, Xs, ls, os_, noise):
    """R&W eqs. 2.25-2.26 for one output with a zero mean function, via Cholesky (their Algorithm 2.1)."""
    K = rbf(X, X, ls, os_) + noise * np.eye(len(X))
    cf = scipy.linalg.cho_factor(K, lower=True)
    ks = rbf(X, Xs, ls, os_)
    mean = ks.T @ scipy.linalg.cho_solve(cf, y)
    v = scipy.linalg.solve_triangular(cf[0], ks, lower=True)
    var = os_ - (v ** 2).sum(0)
    return mean, var


def rw_log_marginal(X, y, ls, os_, noise):
    """R&W eq. 2.30."""
    K = rbf(X, X, ls, os_) + noise * np.eye(len(X))
    cf = scipy.linalg.cho_factor(K, lower=True)
    alpha = scipy.linalg.cho_solve(cf, y)
    return -0.5 * y @ alpha - np.log(np.diag(cf[0])).sum() - 0.5 * len(X) * np.log(2 * np.pi)


def rw_log_marginal_grad(X, y, ls, os_, noise):
    """R&W eq. 5.9 w.r.t. (lengthscale, outputscale, noise)."""
    n = len(X)
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    Kf = os_ * np.exp(-0.5 * d2 / ls ** 2)
    K = Kf + noise * np.eye(n)
    Kinv = np.linalg.inv(K)
    alpha = Kinv @ y
    W = np.outer(alpha, alpha) - Kinv
    return np.array([0.5 * np.trace(W @ (Kf * d2 / ls ** 3)), 0.5 * np.trace(W @ (Kf / os_)), 0.5 * np.trace(W)])


class _ZeroMeanBatchGP(gpytorch.models.ExactGP):
    """MixingBatchGP (online_model.py:286-315) with a zero nominal model: independent outputs, ScaleKernel(RBF) per output."""

    def __init__(self, X, Y, likelihood):
        super().__init__(X, Y, likelihood)
        ny = Y.shape[1]
        self.covar_module = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(batch_shape=torch.Size([ny])),
                                                         batch_shape=torch.Size([ny]))
        self.ny = ny

    def forward(self, x):
        mean = torch.zeros(self.ny, x.shape[0], dtype=x.dtype)
        return gpytorch.distributions.MultitaskMultivariateNormal.from_batch_mvn(
            gpytorch.distributions.MultivariateNormal(mean, self.covar_module(x)))


@pytest.fixture(scope='module')
def problem():
    rng = np.random.RandomState(4)
    n, d, ny = 23, 4, 3
    X = rng.randn(n, d) * 0.7
    Y = np.stack([np.sin(X @ rng.randn(d)) + 0.05 * rng.randn(n) for _ in range(ny)], axis=1)
    Xs = rng.randn(9, d) * 0.8
    raw = dict(ls=rng.randn(ny) * 0.5, os=rng.randn(ny) * 0.5, tn=rng.randn(ny) * 0.5 - 2.0, gn=np.array([-2.5]))
    lik = gpytorch.likelihoods.MultitaskGaussianLikelihood(num_tasks=ny).double()
    gp = _ZeroMeanBatchGP(torch.from_numpy(X), torch.from_numpy(Y), lik).double()
    with torch.no_grad():
        gp.covar_module.base_kernel.raw_lengthscale.copy_(torch.from_numpy(raw['ls']).view(ny, 1, 1))
        gp.covar_module.raw_outputscale.copy_(torch.from_numpy(raw['os']))
        lik.raw_task_noises.copy_(torch.from_numpy(raw['tn']))
        lik.raw_noise.copy_(torch.from_numpy(raw['gn']))
    hyp = dict(ls=softplus(raw['ls']), os=softplus(raw['os']), noise=softplus(raw['tn']) + 1e-4 + softplus(raw['gn']) + 1e-4)
    return X, Y, Xs, gp, lik, hyp


def test_shim_parameterisation_is_gpytorchs(problem):
    X, Y, Xs, gp, lik, hyp = problem
    np.testing.assert_allclose(gp.covar_module.base_kernel.lengthscale.detach().numpy().reshape(-1), hyp['ls'], rtol=1e-14)
    np.testing.assert_allclose(gp.covar_module.outputscale.detach().numpy(), hyp['os'], rtol=1e-14)
    np.testing.assert_allclose((lik.task_noises + lik.noise).detach().numpy(), hyp['noise'], rtol=1e-14)


def test_shim_posterior_is_rasmussen_williams_2_25_and_2_26(problem):
    X, Y, Xs, gp, lik, hyp = problem
    gp.eval()
    lik.eval()
    with torch.no_grad(), gpytorch.settings.fast_pred_var():
        pred = lik(gp(torch.from_numpy(Xs)))
        mean, var = pred.mean.numpy(), pred.variance.numpy()
    for d in range(Y.shape[1]):
        m, v = rw_posterior(X, Y[:, d], Xs, hyp['ls'][d], hyp['os'][d], hyp['noise'][d])
        np.testing.assert_allclose(mean[:, d], m, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(var[:, d], v + hyp['noise'][d], rtol=1e-10, atol=1e-12)  # likelihood(.) adds the task noise


def test_shim_marginal_likelihood_is_rasmussen_williams_2_30_and_its_gradient_5_9(problem):
    X, Y, Xs, gp, lik, hyp = problem
    gp.train()
    lik.train()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, gp)
    out = gp(torch.from_numpy(X))
    loss = -mll(out, torch.from_numpy(Y))
    want = sum(rw_log_marginal(X, Y[:, d], hyp['ls'][d], hyp['os'][d], hyp['noise'][d]) for d in range(Y.shape[1]))
    # gpytorch 1.1.1's ExactMarginalLogLikelihood.forward divides by target.size(-1) — for (N, ny) multitask targets that is
    # the number of TASKS, not of data points; a constant factor on the loss that Adam's step is invariant to only in
    # direction, so the shim (and the device fit) keep it
    scale = Y.shape[1]
    assert abs(float(loss.detach()) + want / scale) < 1e-10 * abs(want)
    loss.backward()
    g_ls = gp.covar_module.base_kernel.raw_lengthscale.grad.numpy().reshape(-1)
    g_os = gp.covar_module.raw_outputscale.grad.numpy()
    g_tn = lik.raw_task_noises.grad.numpy()
    g_gn = float(lik.raw_noise.grad)
    sig = lambda raw_value: 1.0 / (1.0 + np.exp(-raw_value))  # d softplus / d raw
    raw_ls = gp.covar_module.base_kernel.raw_lengthscale.detach().numpy().reshape(-1)
    raw_os = gp.covar_module.raw_outputscale.detach().numpy()
    raw_tn = lik.raw_task_noises.detach().numpy()
    raw_gn = float(lik.raw_noise.detach())
    total_gn = 0.0
    for d in range(Y.shape[1]):
        g = rw_log_marginal_grad(X, Y[:, d], hyp['ls'][d], hyp['os'][d], hyp['noise'][d]) / scale
        np.testing.assert_allclose(g_ls[d], -g[0] * sig(raw_ls[d]), rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(g_os[d], -g[1] * sig(raw_os[d]), rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(g_tn[d], -g[2] * sig(raw_tn[d]), rtol=1e-8, atol=1e-12)
        total_gn += -g[2] * sig(raw_gn)
    np.testing.assert_allclose(g_gn, total_gn, rtol=1e-8, atol=1e-12)
    # ... and eq. 5.9 itself against central differences of eq. 2.30
    d, h = 1, 1e-6
    args = [hyp['ls'][d], hyp['os'][d], hyp['noise'][d]]
    ana = rw_log_marginal_grad(X, Y[:, d], *args)
    for i in range(3):
        hi, lo = list(args), list(args)
        hi[i] += h
        lo[i] -= h
        fd = (rw_log_marginal(X, Y[:, d], *hi) - rw_log_marginal(X, Y[:, d], *lo)) / (2 * h)
        assert abs(fd - ana[i]) < 1e-5 * max(1.0, abs(ana[i]))


def test_oracle_gp_posterior_is_rasmussen_williams(problem):
    """oracle.mppi_oracle.gp_posterior (what the device kernels are compared with): residual targets, per-output
    hyper-parameters, variance INCLUDING the likelihood noise (online_model.py:435-449 samples likelihood(gp(xu)))."""
    X, Y, Xs, gp, lik, hyp = problem
    g = S.GPSpec(X, Y, hyp['ls'], hyp['os'], hyp['noise'], sample=True)
    d = S.DynamicsSpec(kind=S.DYN_MLP, nx=3, nu=1, net=S.MLPSpec([np.zeros((3, 4))], [np.zeros(3)]), gp=g)
    mean, var = O.gp_posterior(d, torch.from_numpy(Xs))
    for k in range(Y.shape[1]):
        m, v = rw_posterior(X, Y[:, k], Xs, hyp['ls'][k], hyp['os'][k], hyp['noise'][k])
        np.testing.assert_allclose(mean[:, k].numpy(), m, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(var[:, k].numpy(), v + hyp['noise'][k], rtol=1e-10, atol=1e-12)
