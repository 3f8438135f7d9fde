"""Import the UNMODIFIED reference (`/root/reference/tampc`) over the shim and build its real controllers.

TEST INFRASTRUCTURE ONLY — usable only where /root/reference exists (this container, not the GPU box).
Used by oracle/make_golden.py to generate tests/golden/*.npz and by the `not gpu` tests that pin the
oracle restatement (oracle/mppi_oracle.py) to the reference's own in-tree code.

Restated here (the scripts cannot be imported: they parse argv and open log files at module scope,
push_main.py:39,1758; peg_main.py:35,712): the getter pieces `ds()`, `pre_invariant_preprocessor()`,
`controller_options()` of push_main.py:57-113 and peg_main.py:45-106, and `constrain_state`
(push_main.py:190-193).  Everything else — datasets, transforms, checkpoints, dynamics, costs, MPPI —
is the reference's own code."""
import math
import os
import sys
import warnings

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get('TAMPC_REFERENCE_ROOT', '/root/reference')
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'tampc'))


_ref = None


def load_reference():
    """Returns a namespace of the reference modules (imported once)."""
    global _ref
    if _ref is not None:
        return _ref
    if not available():
        raise RuntimeError('reference tree not present at {}'.format(REFERENCE_ROOT))
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import _stubs
    _stubs.install()
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(1, REFERENCE_ROOT)
    warnings.filterwarnings('ignore', category=SyntaxWarning)
    import types
    ns = types.SimpleNamespace()
    from tampc.controller import controller, online_controller, gating_function
    from tampc.dynamics import model, hybrid_model, prior, online_model
    from tampc.transform import invariant
    from tampc import cost, util, cfg
    from tampc.env import block_push, peg_in_hole, env as env_mod
    from arm_pytorch_utilities import preprocess, load_data, math_utils
    ns.controller, ns.online_controller, ns.gating_function = controller, online_controller, gating_function
    ns.model, ns.hybrid_model, ns.prior, ns.invariant = model, hybrid_model, prior, invariant
    ns.online_model = online_model
    ns.cost, ns.util, ns.cfg = cost, util, cfg
    ns.block_push, ns.peg_in_hole, ns.env_mod = block_push, peg_in_hole, env_mod
    ns.preprocess, ns.load_data, ns.math_utils = preprocess, load_data, math_utils
    _ref = ns
    return ns


# ----------------------------------------------------------------------------------------------
# task pieces restated from the scripts

def _push_pre_invariant(ref):
    # push_main.py:66-73 (REX_EXTRACT branch)
    return ref.preprocess.PytorchTransformer(ref.preprocess.NullSingleTransformer(),
                                             ref.preprocess.RobustMinMaxScaler(feature_range=[[0, 0, 0], [3, 3, 1.5]]))


def _peg_pre_invariant(ref):
    # peg_main.py:60-67
    return ref.preprocess.PytorchTransformer(ref.preprocess.NullSingleTransformer(),
                                             ref.preprocess.RobustMinMaxScaler())


def push_constrain_state(state):
    # push_main.py:190-193
    from arm_pytorch_utilities import math_utils
    state[:, 2] = math_utils.angle_normalize(state[:, 2])
    return state


def push_options(ref):
    """push_main.py:76-113."""
    env_cls = ref.block_push.PushPhysicallyAnyAlongEnv
    u_min, u_max = env_cls.get_control_bounds()
    common = {
        'Q': torch.tensor(env_cls.state_cost(), dtype=torch.double), 'R': 0.01, 'R_env': env_cls.control_cost(),
        'recovery_scale': 2000, 'u_min': u_min, 'u_max': u_max, 'compare_to_goal': env_cls.state_difference,
        'state_dist': env_cls.state_distance, 'u_similarity': env_cls.control_similarity, 'device': 'cpu',
        'terminal_cost_multiplier': 50, 'abs_unrecognized_threshold': 10 / 1.1407,
    }
    mpc = {
        'num_samples': 500, 'noise_sigma': torch.diag(torch.tensor([0.2, 0.4, 0.7], dtype=torch.double)),
        'noise_mu': torch.tensor([0, 0.1, 0], dtype=torch.double), 'lambda_': 1e-2, 'horizon': 40,
        'u_init': torch.tensor([0, 0.5, 0], dtype=torch.double), 'sample_null_action': False,
        'step_dependent_dynamics': True, 'rollout_samples': 10, 'rollout_var_cost': 0,
    }
    return common, mpc


def peg_options(ref):
    """peg_main.py:70-106."""
    env_cls = ref.peg_in_hole.PegFloatingGripperEnv
    u_min, u_max = env_cls.get_control_bounds()
    common = {
        'Q': torch.tensor(env_cls.state_cost(), dtype=torch.double), 'R': 0.01, 'u_min': u_min, 'u_max': u_max,
        'compare_to_goal': env_cls.state_difference, 'state_dist': env_cls.state_distance,
        'u_similarity': env_cls.control_similarity, 'device': 'cpu', 'terminal_cost_multiplier': 50,
        'trap_cost_annealing_rate': 0.9, 'abs_unrecognized_threshold': 15 / 1.2185,
    }
    mpc = {
        'num_samples': 500, 'noise_sigma': torch.diag(torch.tensor([0.2, 0.2], dtype=torch.double)),
        'noise_mu': torch.tensor([0, 0], dtype=torch.double), 'lambda_': 1e-2, 'horizon': 10,
        'u_init': torch.tensor([0, 0], dtype=torch.double), 'sample_null_action': False,
        'step_dependent_dynamics': True, 'rollout_samples': 10, 'rollout_var_cost': 0,
    }
    return common, mpc


def _fake_env(cls):
    """An env instance without running its pybullet constructor; only type(env) and static fns are used."""
    return object.__new__(cls)


def build_task(task, num_samples=None, rollout_samples=None, horizon=None, nominal_adapt='NONE', controller_class=None):
    """Build the reference's real OnlineMPPI controller for 'push' or 'peg' exactly as
    run_controller does (push_main.py:796-892 / peg_main.py:221-367), REX_EXTRACT transform,
    AlwaysSelectNominal gating.  nominal_adapt: name of a hybrid_model.OnlineAdapt member ('NONE' is the scripts'
    default; 'GP_KERNEL_INDEP_OUT' is --always_estimate_error, push_main.py:1788-1789)."""
    ref = load_reference()
    UseTsf = ref.util.UseTsf
    if task == 'push':
        env = _fake_env(ref.block_push.PushPhysicallyAnyAlongEnv)
        config = ref.load_data.DataConfig(predict_difference=True, predict_all_dims=True, expanded_input=False)
        ds = ref.block_push.PushDataSource(env, data_dir='pushing/physical0.mat', config=config, device='cpu',
                                           validation_ratio=0.1)
        pre = lambda use_tsf: _push_pre_invariant(ref)
        rep_name, prefix = 'saved', 'dynamics'
        common, mpc = push_options(ref)
        constrain = push_constrain_state
    elif task == 'peg':
        env = _fake_env(ref.peg_in_hole.PegFloatingGripperEnv)
        config = ref.load_data.DataConfig(predict_difference=True, predict_all_dims=True, expanded_input=False)
        ds = ref.peg_in_hole.PegInHoleDataSource(env, data_dir='peg/floating0.mat', config=config, device='cpu',
                                                 validation_ratio=0.1)
        pre = lambda use_tsf: _peg_pre_invariant(ref)
        rep_name, prefix = 'saved_peg', 'peg'
        common, mpc = peg_options(ref)
        constrain = ref.online_controller.noop_constrain
    else:
        raise ValueError(task)
    with torch.no_grad():
        untransformed_config, tsf_name, preprocessor = ref.util.update_ds_with_transform(
            env, ds, UseTsf.REX_EXTRACT, pre, evaluate_transform=False, rep_name=rep_name)

        class _Getter(ref.util.EnvGetter):
            @staticmethod
            def dynamics_prefix():
                return prefix

        pm = _Getter.loaded_prior(ref.prior.NNPrior, ds, tsf_name, False)
        with torch.enable_grad():  # OnlineGPMixing fits its hyper-parameters by Adam in its constructor
            hybrid = ref.hybrid_model.HybridDynamicsModel(
                [ds], pm, env.state_difference, [tsf_name], device='cpu', preprocessor=ref.util.no_tsf_preprocessor(),
                nominal_model_kwargs={'online_adapt': getattr(ref.hybrid_model.OnlineAdapt, nominal_adapt)},
                local_model_kwargs={})
        if num_samples is not None:
            mpc['num_samples'] = num_samples
        if rollout_samples is not None:
            mpc['rollout_samples'] = rollout_samples
        if horizon is not None:
            mpc['horizon'] = horizon
        ctrl = (controller_class or ref.online_controller.OnlineMPPI)(ds, hybrid, ds.original_config(),
                                                gating=ref.gating_function.AlwaysSelectNominal(),
                                                autonomous_recovery=ref.online_controller.AutonomousRecovery.MAB,
                                                constrain_state=constrain, mpc_opts=mpc, **common)
    return ctrl, ds, env


class inject_gp_draws:
    """Context manager: every `torch.distributions.Normal.sample()` made by OnlineGPMixing.sample
    (online_model.py:446-449) returns loc + scale * eps with eps popped from `draws` (a list of (rows, ny) tensors, one
    per horizon step); with the list exhausted it returns loc (a zero draw)."""

    def __init__(self, draws):
        self.draws = list(draws)

    def __enter__(self):
        import torch.distributions as D
        self._orig = D.Normal.sample
        outer = self

        def sample(self_, sample_shape=torch.Size()):
            if outer.draws and tuple(outer.draws[0].shape) == tuple(self_.loc.shape):
                return self_.loc + self_.scale * outer.draws.pop(0)
            return self_.loc.clone()

        D.Normal.sample = sample
        return self

    def __exit__(self, *exc):
        import torch.distributions as D
        D.Normal.sample = self._orig
        return False
