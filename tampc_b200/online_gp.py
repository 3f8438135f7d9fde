"""Drop-in for `tampc.dynamics.online_model.OnlineGPMixing` (independent outputs) whose hyper-parameter fit runs on the
device: the reference refits its GP with 15 Adam iterations of autograd through gpytorch on every
control step spent in non-nominal dynamics (online_model.py:386-430; 150 at a reset) — the other heavy per-step cost
next to the rollout.  Here the data bookkeeping of `OnlineGPMixing` is kept verbatim in spirit (roll / append, the three
refit strategies, optimiser state carried across updates) and `_fit_params` becomes one call of `tampc_gp_fit`.

The class is built against the reference's own `OnlineDynamicsModel` base, so `update()` (state difference,
preprocessor transform; online_model.py:33-44) is the reference's code and `HybridDynamicsModel._uses_local_model_api`
recognises instances:

    from tampc.dynamics import online_model
    from tampc_b200.online_gp import install_gp
    install_gp(online_model)          # online_model.OnlineGPMixing = the device-fitted class; no reference file edited

`tampc_b200.extract` reads the fitted object like the original (duck-typed `.gp.covar_module`, `.likelihood`), so the
rollout's GP residual (`tampc_mppi_set_gp`) follows it on every command.  Prediction is not offered on the host: the
rollout, `predict_next_state` and the one-step sampler all evaluate the GP inside the kernels (no CPU fallback).
"""
import ctypes as C
import types

import numpy as np
import torch

from . import _lib as L


def _softplus(x):
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


class _RBFKernelView:
    def __init__(self, owner):
        self._o = owner

    @property
    def lengthscale(self):
        return torch.from_numpy(_softplus(self._o.raw_lengthscale)).view(-1, 1, 1)


class _ScaleKernelView:
    def __init__(self, owner):
        self._o = owner
        self.base_kernel = _RBFKernelView(owner)

    @property
    def outputscale(self):
        return torch.from_numpy(_softplus(self._o.raw_outputscale))


class _LikelihoodView:
    def __init__(self, owner):
        self._o = owner

    @property
    def task_noises(self):
        return torch.from_numpy(_softplus(self._o.raw_task_noises) + 1e-4)

    @property
    def noise(self):
        return torch.from_numpy(_softplus(self._o.raw_noise) + 1e-4)


class _GPView:
    """What `extract.gp_from_online_model` reads of a gpytorch ExactGP."""

    def __init__(self, owner):
        self._o = owner
        self.covar_module = _ScaleKernelView(owner)
        self.training = True  # eval() only after a fit, like the reference (online_model.py:426-427)

    def forward(self, xu):
        return types.SimpleNamespace(mean=self._o._mean_function(xu))


def make_online_gp_class(online_model):
    """`online_model`: the reference module `tampc.dynamics.online_model` (anything exposing `OnlineDynamicsModel`)."""
    Base = online_model.OnlineDynamicsModel
    strategies = getattr(online_model, 'RefitGPStrategy', types.SimpleNamespace(RESET_DATA=0, RECREATE_GP=1, RECREATE_ALL=2))

    class B200OnlineGP(Base):
        """OnlineGPMixing (online_model.py:324-457) with `_fit_params` on the device."""

        def __init__(self, prior, ds, state_difference, max_data_points=50, training_iter=100, slice_to_use=None,
                     refit_strategy=strategies.RESET_DATA, use_independent_outputs=True, allow_update=True, sample=True,
                     cuda_device=0, **kwargs):
            if not use_independent_outputs:
                raise NotImplementedError("only the independent-output GP (OnlineAdapt.GP_KERNEL_INDEP_OUT) is fused")
            if max_data_points > L.GP_FIT_MAX_POINTS:
                raise ValueError("max_data_points {} exceeds TAMPC_GP_FIT_MAX_POINTS={}".format(max_data_points, L.GP_FIT_MAX_POINTS))
            super().__init__(ds, state_difference, **kwargs)
            self._lib = L.load()
            self.cuda_device = cuda_device
            self.prior = prior
            self.max_data_points = max_data_points
            self.xu = None
            self.y = None
            self.last_loss = 0
            self.last_loss_diff = 0
            self.training_iter = training_iter
            self.refit_strategy = refit_strategy
            self.use_independent_outputs = True
            self.allow_update = allow_update
            self.sample_dynamics = sample
            self.last_prediction = None
            if slice_to_use is None:
                slice_to_use = slice(max_data_points)
            xu, y, _ = self.ds.training_set()
            self.init_xu = xu[slice_to_use]
            self.init_y = y[slice_to_use]
            self.ny = int(self.ds.config.ny)
            self.reset()

        # ---- the pieces of gpytorch state the reference keeps (online_model.py:363-384)
        def _recreate_all(self):
            self.raw_task_noises = np.zeros(self.ny)   # MultitaskGaussianLikelihood: fresh raw parameters are 0
            self.raw_noise = np.zeros(1)
            self.likelihood = _LikelihoodView(self)
            self._recreate_gp()

        def _recreate_gp(self):
            self.raw_lengthscale = np.zeros(self.ny)   # ScaleKernel(RBFKernel(batch_shape=[ny]))
            self.raw_outputscale = np.zeros(self.ny)
            self.gp = _GPView(self)
            # MixingBatchGP.__init__ (online_model.py:289-295): is the nominal model evaluated in the original space?
            try:
                self.ds.preprocessor.invert_x(self.xu)
                self.nominal_in_orig_space = True
            except RuntimeError:
                self.nominal_in_orig_space = False
            n = 3 * self.ny + 1                           # a new torch.optim.Adam: zero moments, step 0
            self._adam_m, self._adam_v, self._adam_step = np.zeros(n), np.zeros(n), C.c_int64(0)

        def reset(self):
            self.xu = self.init_xu.clone()
            self.y = self.init_y.clone()
            self._recreate_all()
            if len(self.xu):
                self._fit_params(self.training_iter)

        def _mean_function(self, xu):
            """MixingBatchGP.forward's mean (online_model.py:304-311): the nominal model at the GP's inputs, (N, ny)."""
            pre = self.ds.preprocessor
            with torch.no_grad():
                if self.nominal_in_orig_space:
                    pred = self.prior.get_batch_predictions(pre.invert_x(xu), all_in_transformed_space=False)
                    return pre.transform_y(pred)
                return self.prior.get_batch_predictions(xu, all_in_transformed_space=True)

        def _update(self, px, pu, y):
            # online_model.py:386-410
            self.last_prediction = None
            if not self.allow_update:
                return
            xu = torch.cat((px.view(1, -1), pu.view(1, -1)), dim=1)
            y = y.view(1, -1)
            if self.xu.shape[0] < self.max_data_points:
                self.xu = torch.cat((self.xu, xu), dim=0)
                self.y = torch.cat((self.y, y), dim=0)
            else:
                self.xu = torch.roll(self.xu, -1, dims=0)
                self.xu[-1] = xu
                self.y = torch.roll(self.y, -1, dims=0)
                self.y[-1] = y
            if self.refit_strategy is strategies.RESET_DATA or self.refit_strategy == strategies.RESET_DATA:
                self._fit_params(self.training_iter // 10)
            elif self.refit_strategy == strategies.RECREATE_GP:
                self._recreate_gp()
                self._fit_params(self.training_iter // 3)
            else:
                self._recreate_all()
                self._fit_params(self.training_iter)

        def _fit_params(self, training_iter):
            """online_model.py:412-430 as one device call."""
            X = np.ascontiguousarray(self.xu.detach().to('cpu', torch.double).numpy())
            resid = np.ascontiguousarray((self.y - self._mean_function(self.xu)).detach().to('cpu', torch.double).numpy())
            raw = np.concatenate([self.raw_lengthscale, self.raw_outputscale, self.raw_task_noises, self.raw_noise])
            loss = np.zeros(2)
            rc = self._lib.tampc_gp_fit(self.cuda_device, X.shape[0], X.shape[1], self.ny, L.dptr(X), L.dptr(resid), int(training_iter),
                                        0.1, L.dptr(raw), L.dptr(self._adam_m), L.dptr(self._adam_v), C.byref(self._adam_step),
                                        L.dptr(loss))
            if rc != L.OK:
                raise RuntimeError("tampc_gp_fit error {}: {}".format(rc, self._lib.tampc_gp_fit_last_error().decode()))
            ny = self.ny
            self.raw_lengthscale, self.raw_outputscale = raw[:ny].copy(), raw[ny:2 * ny].copy()
            self.raw_task_noises, self.raw_noise = raw[2 * ny:3 * ny].copy(), raw[3 * ny:].copy()
            if training_iter >= 2:
                self.last_loss_diff = float(loss[0] - loss[1])
            elif training_iter == 1:
                self.last_loss_diff = float(loss[0]) - self.last_loss
            if training_iter >= 1:
                self.last_loss = float(loss[0])
            self.gp.training = False

        def _dynamics_in_transformed_space(self, px, pu, cx, cu):
            raise RuntimeError("B200OnlineGP predicts inside the rollout kernels only (tampc_mppi_set_gp): use the installed "
                               "B200MPPI's command / predict_next_state / apply_dynamics; there is no host prediction path")

        def predict(self, px, pu, cx, cu, already_transformed=False):
            return self._dynamics_in_transformed_space(px, pu, cx, cu)

    return B200OnlineGP


def install_gp(online_model, **defaults):
    """Point the reference module's `OnlineGPMixing` at the device-fitted class (the reference's
    `HybridDynamicsModel.get_local_model` looks the name up at call time, hybrid_model.py:142-146)."""
    cls = make_online_gp_class(online_model)
    online_model.OnlineGPMixing = cls
    return cls
