"""Drop-in for `tampc.dynamics.online_model.OnlineGPMixing` (independent outputs) whose hyper-parameter fit runs on the
device: the reference refits its GP with 15 Adam iterations of autograd through gpytorch on every
control step spent in non-nominal dynamics (online_model.py:386-430; 150 at a reset) — the other heavy per-step cost
next to the rollout.  The class here SUBCLASSES the reference's own `OnlineGPMixing`: its constructor, `reset`, `update`
/ `_update` (roll / append of the data window, the three refit strategies) run unchanged; only the gpytorch objects
become plain raw-parameter holders and `_fit_params` becomes one call of `tampc_gp_fit` (optimiser state carried across
updates like the reference's Adam instance).  `HybridDynamicsModel._uses_local_model_api` recognises instances:

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

    def set_train_data(self, inputs=None, targets=None, strict=True):
        """gpytorch ExactGP.set_train_data as the reference's `_update` calls it (online_model.py:401): the data live on
        the owner (`xu`, `y`), which `_fit_params` reads."""


def make_online_gp_class(online_model):
    """`online_model`: the reference module `tampc.dynamics.online_model`.  The returned class SUBCLASSES its
    `OnlineGPMixing`: construction, `reset`, `update` / `_update` (data window, refit strategies) stay the reference's own
    code; only what touches gpytorch is replaced — the likelihood / GP objects (plain raw-parameter holders here) and
    `_fit_params` (one `tampc_gp_fit` call)."""
    Base = online_model.OnlineGPMixing

    class B200OnlineGP(Base):
        def __init__(self, prior, ds, state_difference, *args, cuda_device=0, **kwargs):
            kwargs.setdefault('use_independent_outputs', True)  # OnlineAdapt.GP_KERNEL_INDEP_OUT (push_main.py:1789)
            if not kwargs['use_independent_outputs']:
                raise NotImplementedError("only the independent-output GP (OnlineAdapt.GP_KERNEL_INDEP_OUT) is fused")
            max_points = kwargs.get('max_data_points', args[0] if args else 50)
            if max_points > L.GP_FIT_MAX_POINTS:
                raise ValueError("max_data_points {} exceeds TAMPC_GP_FIT_MAX_POINTS={}".format(max_points, L.GP_FIT_MAX_POINTS))
            # needed by the overrides below, which the reference constructor reaches through reset()
            self._lib = L.load()
            self.cuda_device = cuda_device
            self.ny = int(ds.config.ny)
            super().__init__(prior, ds, state_difference, *args, **kwargs)

        # ---- the gpytorch objects of online_model.py:363-384 as raw-parameter holders
        def _recreate_all(self):
            self.raw_task_noises = np.zeros(self.ny)   # MultitaskGaussianLikelihood: fresh raw parameters are 0
            self.raw_noise = np.zeros(1)
            self.likelihood = _LikelihoodView(self)
            self._recreate_gp()

        def _recreate_gp(self):
            self.raw_lengthscale = np.zeros(self.ny)   # ScaleKernel(RBFKernel(batch_shape=[ny]))
            self.raw_outputscale = np.zeros(self.ny)
            self.gp = _GPView(self)
            self.optimizer = None
            # MixingBatchGP.__init__ (online_model.py:289-295): is the nominal model evaluated in the original space?
            try:
                self.ds.preprocessor.invert_x(self.xu)
                self.nominal_in_orig_space = True
            except RuntimeError:
                self.nominal_in_orig_space = False
            n = 3 * self.ny + 1                           # a new torch.optim.Adam: zero moments, step 0
            self._adam_m, self._adam_v, self._adam_step = np.zeros(n), np.zeros(n), C.c_int64(0)

        def _mean_function(self, xu):
            """MixingBatchGP.forward's mean (online_model.py:304-311): the nominal model at the GP's inputs, (N, ny)."""
            pre = self.ds.preprocessor
            with torch.no_grad():
                if self.nominal_in_orig_space:
                    pred = self.prior.get_batch_predictions(pre.invert_x(xu), all_in_transformed_space=False)
                    return pre.transform_y(pred)
                return self.prior.get_batch_predictions(xu, all_in_transformed_space=True)

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

        def _make_prediction(self, cx, cu):
            raise RuntimeError("B200OnlineGP predicts inside the rollout kernels only (tampc_mppi_set_gp): use the installed "
                               "B200MPPI's command / predict_next_state / apply_dynamics; there is no host prediction path")

        def predict(self, px, pu, cx, cu, already_transformed=False):
            return self._make_prediction(cx, cu)

    return B200OnlineGP


def install_gp(online_model, **defaults):
    """Point the reference module's `OnlineGPMixing` at the device-fitted class (the reference's
    `HybridDynamicsModel.get_local_model` looks the name up at call time, hybrid_model.py:142-146)."""
    cls = make_online_gp_class(online_model)
    online_model.OnlineGPMixing = cls
    return cls
