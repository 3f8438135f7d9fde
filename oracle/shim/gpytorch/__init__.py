"""gpytorch restated — just the exact-GP pieces tampc's OnlineGPMixing touches (tampc/dynamics/online_model.py:
252-457).  TEST INFRASTRUCTURE ONLY; source of gpytorch 1.1.1 (README.md:57,86 of the reference) is not under
/root/reference and cannot be fetched: PARITY UNPINNED upstream.

Restated from the published exact-GP equations (Rasmussen & Williams 2006, eqs. 2.25-2.26, 2.30) and gpytorch's
documented parameterisation:
  kernels.RBFKernel(batch_shape)     k(a,b) = exp(-|a-b|^2 / (2 l^2)),  l = softplus(raw_lengthscale), raw init 0
  kernels.ScaleKernel(base)          s * base,                          s = softplus(raw_outputscale), raw init 0
  likelihoods.MultitaskGaussianLikelihood(num_tasks)  per-task noise = (softplus(raw_task_noises)+1e-4) +
                                     (softplus(raw_noise)+1e-4)  (rank 0: diagonal task noise + global noise,
                                     both under GreaterThan(1e-4))
  models.ExactGP                     train(): prior at the training inputs; eval(): posterior
                                       mean* = m(x*) + K*^T (K + S)^-1 (y - m(X)),  var* = k** - K*^T (K + S)^-1 K*
                                     likelihood(posterior) adds the per-task noise to var*
  mlls.ExactMarginalLogLikelihood    log N(y | m(X), K + S) / target.size(-1)
  settings.fast_pred_var             (LOVE) exact here: with <= 50 points the Lanczos rank covers the data
MixingFullGP's pieces (MultitaskKernel / MultitaskMean, rank-1 task covariance) are import-only placeholders:
the task scripts use the independent-output GP (hybrid_model.py:125, push_main.py:1789-1791).
"""
import contextlib
import math

import torch
from torch.nn import Parameter
from torch.nn.functional import softplus


class Module(torch.nn.Module):
    pass


class _LazyKernel:
    """Kernel matrix between two point sets, evaluated on demand (the test-test block is never formed)."""

    def __init__(self, kernel, x1, x2):
        self.kernel, self.x1, self.x2 = kernel, x1, x2

    def evaluate(self):
        return self.kernel.forward(self.x1, self.x2)

    def diag(self):
        return self.kernel.forward(self.x1, self.x2, diag=True)

    def block(self, rows, cols):
        return _LazyKernel(self.kernel, self.x1[rows], self.x2[cols])


class _Kernel(Module):
    def __call__(self, x1, x2=None):
        return _LazyKernel(self, x1, x1 if x2 is None else x2)


class _Namespace:
    pass


kernels = _Namespace()
means = _Namespace()
models = _Namespace()
likelihoods = _Namespace()
distributions = _Namespace()
mlls = _Namespace()
settings = _Namespace()


class RBFKernel(_Kernel):
    def __init__(self, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.batch_shape = torch.Size(batch_shape)
        self.raw_lengthscale = Parameter(torch.zeros(*self.batch_shape, 1, 1))

    @property
    def lengthscale(self):
        return softplus(self.raw_lengthscale)

    def forward(self, x1, x2, diag=False):
        if diag:
            return torch.ones(*self.batch_shape, x1.shape[-2], dtype=x1.dtype, device=x1.device)
        a = x1.unsqueeze(-3) if self.batch_shape else x1
        b = x2.unsqueeze(-3) if self.batch_shape else x2
        a = a / self.lengthscale
        b = b / self.lengthscale
        d2 = (a.unsqueeze(-2) - b.unsqueeze(-3)).pow(2).sum(-1)
        return torch.exp(-0.5 * d2)


class ScaleKernel(_Kernel):
    def __init__(self, base_kernel, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.base_kernel = base_kernel
        self.batch_shape = torch.Size(batch_shape)
        self.raw_outputscale = Parameter(torch.zeros(*self.batch_shape))

    @property
    def outputscale(self):
        return softplus(self.raw_outputscale)

    def forward(self, x1, x2, diag=False):
        k = self.base_kernel.forward(x1, x2, diag=diag)
        s = self.outputscale
        return k * (s.view(*s.shape, 1) if diag else s.view(*s.shape, 1, 1))


class _Unsupported(Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("oracle shim: {} (MixingFullGP path) is not restated".format(type(self).__name__))


class MultitaskKernel(_Unsupported):
    pass


class MultitaskMean(_Unsupported):
    pass


class ZeroMean(_Unsupported):
    pass


kernels.RBFKernel, kernels.ScaleKernel, kernels.MultitaskKernel = RBFKernel, ScaleKernel, MultitaskKernel
means.MultitaskMean, means.ZeroMean = MultitaskMean, ZeroMean


class MultivariateNormal:
    """Batch of independent GPs: mean (ny, N), lazy covariance (ny, N, N); `noise` (ny,) once a likelihood was applied."""

    def __init__(self, mean, covariance, noise=None):
        self.mean_b, self.covar, self.noise = mean, covariance, noise


class MultitaskMultivariateNormal:
    """Interleaved view of a batch MVN: mean / variance / stddev are (N, ny)."""

    def __init__(self, batch_mvn, var=None):
        self.base = batch_mvn
        self._var = var  # (ny, N) marginal variances when known (posterior)

    @classmethod
    def from_batch_mvn(cls, batch_mvn, task_dim=-1):
        return cls(batch_mvn)

    @property
    def mean(self):
        return self.base.mean_b.transpose(0, 1)

    @property
    def variance(self):
        v = self._var if self._var is not None else self.base.covar.diag()
        if self.base.noise is not None:
            v = v + self.base.noise.view(-1, 1)
        return v.transpose(0, 1)

    @property
    def stddev(self):
        return self.variance.sqrt()

    def confidence_region(self):
        s2 = self.stddev * 2
        return self.mean - s2, self.mean + s2

    def log_prob(self, target):
        """target (N, ny): sum over the independent tasks of log N(y_d | mean_d, K_d + noise_d I)."""
        K = self.base.covar.evaluate()
        if self.base.noise is not None:
            K = K + torch.diag_embed(self.base.noise.view(-1, 1).expand(-1, K.shape[-1]))
        L = torch.linalg.cholesky(K)
        r = (target.transpose(0, 1) - self.base.mean_b).unsqueeze(-1)
        a = torch.cholesky_solve(r, L)
        n = K.shape[-1]
        quad = (r * a).sum((-2, -1))
        logdet = 2 * torch.diagonal(L, dim1=-2, dim2=-1).log().sum(-1)
        return (-0.5 * (quad + logdet + n * math.log(2 * math.pi))).sum()


distributions.MultivariateNormal = MultivariateNormal
distributions.MultitaskMultivariateNormal = MultitaskMultivariateNormal


class MultitaskGaussianLikelihood(Module):
    def __init__(self, num_tasks, rank=0, **kwargs):
        super().__init__()
        assert rank == 0
        self.num_tasks = num_tasks
        self.raw_task_noises = Parameter(torch.zeros(num_tasks))
        self.raw_noise = Parameter(torch.zeros(1))

    @property
    def task_noises(self):
        return softplus(self.raw_task_noises) + 1e-4

    @property
    def noise(self):
        return softplus(self.raw_noise) + 1e-4

    def total_noise(self):
        return self.task_noises + self.noise

    def __call__(self, dist):
        b = dist.base
        return MultitaskMultivariateNormal(MultivariateNormal(b.mean_b, b.covar, self.total_noise()), dist._var)


likelihoods.MultitaskGaussianLikelihood = MultitaskGaussianLikelihood


class ExactGP(Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.train_inputs = (train_inputs,) if torch.is_tensor(train_inputs) else train_inputs
        self.train_targets = train_targets
        self.likelihood = likelihood

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            self.train_inputs = (inputs,) if torch.is_tensor(inputs) else inputs
        if targets is not None:
            self.train_targets = targets

    def __call__(self, x):
        X = self.train_inputs[0]
        if self.training:
            if not torch.equal(X, x):
                raise RuntimeError("You must train on the training inputs!")
            return self.forward(x)
        n = X.shape[0]
        full = self.forward(torch.cat((X, x), dim=0)).base  # batch MVN over train + test points
        tr, te = slice(0, n), slice(n, None)
        noise = self.likelihood.total_noise()
        Ktt = full.covar.block(tr, tr).evaluate() + torch.diag_embed(noise.view(-1, 1).expand(-1, n))
        Kst = full.covar.block(te, tr).evaluate()  # (ny, B, n)
        L = torch.linalg.cholesky(Ktt)
        resid = (self.train_targets.transpose(0, 1) - full.mean_b[:, tr]).unsqueeze(-1)  # (ny, n, 1)
        alpha = torch.cholesky_solve(resid, L)
        mean = full.mean_b[:, te] + (Kst @ alpha).squeeze(-1)
        w = torch.linalg.solve_triangular(L, Kst.transpose(-1, -2), upper=False)  # (ny, n, B)
        var = full.covar.block(te, te).diag() - w.pow(2).sum(-2)
        post = MultivariateNormal(mean, full.covar.block(te, te))
        return MultitaskMultivariateNormal(post, var)


models.ExactGP = ExactGP


class ExactMarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, output, target):
        return self.likelihood(output).log_prob(target) / target.size(-1)


mlls.ExactMarginalLogLikelihood = ExactMarginalLogLikelihood


@contextlib.contextmanager
def fast_pred_var(state=True):
    yield


settings.fast_pred_var = fast_pred_var
