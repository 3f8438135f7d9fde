"""ctypes binding of libtampc_mppi.so (include/tampc_mppi.h).  Structures mirror the header field by field.

The library is the product's compute path; if it is missing or cannot be loaded this module raises — there is
no Python / CPU fallback (BASELINE.json north_star)."""
import ctypes as C
import os

import numpy as np

from . import spec as S

MAX_NX, MAX_NU, MAX_TRAPS, MAX_GOAL_SETS, MAX_SET_GOALS, MAX_LAYERS, MAX_HORIZON = 8, 4, 64, 2, 8, 4, 128
GP_MAX_POINTS, GP_MAX_DIM, MAX_ROLLOUT_SAMPLES, GP_FIT_MAX_POINTS = 64, 8, 32, 50
ABI_VERSION = 1

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_STATE, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5
PREC_F64, PREC_MIXED, PREC_F32, PREC_TF32X3, PREC_F16X3 = 0, 1, 2, 3, 4
PRECISIONS = {'f64': PREC_F64, 'mixed': PREC_MIXED, 'f32': PREC_F32, 'tf32x3': PREC_TF32X3, 'f16x3': PREC_F16X3}
NOISE_PHILOX, NOISE_INJECTED, NOISE_ACTIONS = 0, 1, 2
Q_COST_TOTAL, Q_PERTURBED, Q_NOMINAL, Q_OMEGA, Q_STATS, Q_STATES, Q_NOISE = 0, 1, 2, 3, 4, 5, 6

EXPORTS = [
    'tampc_mppi_last_error', 'tampc_mppi_abi_version', 'tampc_mppi_create', 'tampc_mppi_destroy',
    'tampc_mppi_set_model', 'tampc_mppi_set_cost', 'tampc_mppi_set_nominal', 'tampc_mppi_get_nominal',
    'tampc_mppi_command', 'tampc_mppi_command_begin', 'tampc_mppi_command_finish', 'tampc_mppi_rollout_costs',
    'tampc_mppi_get_rollouts', 'tampc_mppi_query', 'tampc_mppi_select_action', 'tampc_mppi_state_costs', 'tampc_mppi_kernel_name', 'tampc_mppi_debug_stamps', 'tampc_mppi_device_costs', 'tampc_mppi_synchronize',
    'tampc_mppi_command_resident', 'tampc_mppi_set_state_resident', 'tampc_mppi_rollout_resident', 'tampc_mppi_comm_init', 'tampc_mppi_comm_connect', 'tampc_mppi_comm_disconnect', 'tampc_mppi_predict', 'tampc_mppi_predict_nominal',
    'tampc_mppi_set_gp', 'tampc_mppi_set_gp_draws', 'tampc_gp_fit', 'tampc_gp_fit_last_error',
]


class MlpDesc(C.Structure):
    _fields_ = [('n_layers', C.c_int32), ('dims', C.c_int32 * (MAX_LAYERS + 1)), ('w_off', C.c_int32 * MAX_LAYERS),
                ('b_off', C.c_int32 * MAX_LAYERS)]


class AffineDesc(C.Structure):
    _fields_ = [('scale_off', C.c_int32), ('offset_off', C.c_int32)]


class ModelDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('nx', C.c_int32), ('nu', C.c_int32), ('nv', C.c_int32),
                ('leaky_slope', C.c_double), ('net', MlpDesc), ('encoder', MlpDesc), ('decoder', MlpDesc),
                ('ext_w_off', C.c_int32), ('ext_b_off', C.c_int32), ('ne', C.c_int32),
                ('aff_in', AffineDesc), ('aff_out', AffineDesc), ('aff_y', AffineDesc),
                ('wrap_mask', C.c_uint32), ('has_bad_state_guard', C.c_int32), ('bad_state_norm', C.c_double),
                ('pendulum', C.c_double * 6)]


class GpDesc(C.Structure):
    _fields_ = [('n_points', C.c_int32), ('n_in', C.c_int32), ('n_out', C.c_int32), ('space', C.c_int32),
                ('sample', C.c_int32), ('x_off', C.c_int32), ('resid_off', C.c_int32),
                ('lengthscale', C.c_double * GP_MAX_DIM), ('outputscale', C.c_double * GP_MAX_DIM),
                ('noise', C.c_double * GP_MAX_DIM), ('in_scale', C.c_double * GP_MAX_DIM),
                ('in_offset', C.c_double * GP_MAX_DIM), ('out_scale', C.c_double * GP_MAX_DIM),
                ('rollout_samples', C.c_int32), ('rollout_var_cost', C.c_double), ('rollout_var_discount', C.c_double)]


class CostDesc(C.Structure):
    _fields_ = [('kind', C.c_int32), ('nx', C.c_int32), ('nu', C.c_int32), ('ang_mask', C.c_uint32),
                ('dist_scale', C.c_double * MAX_NX), ('usim_mask', C.c_uint32),
                ('Q', C.c_double * (MAX_NX * MAX_NX)), ('R', C.c_double * (MAX_NU * MAX_NU)),
                ('goal', C.c_double * MAX_NX), ('use_trap_cost', C.c_int32), ('n_traps', C.c_int32),
                ('trap_x', (C.c_double * MAX_NX) * MAX_TRAPS), ('trap_u', (C.c_double * MAX_NU) * MAX_TRAPS),
                ('trap_weight', C.c_double), ('has_terminal', C.c_int32), ('terminal_multiplier', C.c_double),
                ('n_sets', C.c_int32), ('set_size', C.c_int32 * MAX_GOAL_SETS),
                ('set_goals', ((C.c_double * MAX_NX) * MAX_SET_GOALS) * MAX_GOAL_SETS),
                ('set_weight', C.c_double * MAX_GOAL_SETS), ('recovery_scale', C.c_double),
                ('apf_use_goal', C.c_int32), ('apf_helicoid', C.c_int32), ('apf_clockwise', C.c_int32),
                ('apf_max_dist', C.c_double), ('apf_gain', C.c_double), ('apf_rxy', C.c_double * 2), ('apf_oxy', C.c_double * 2)]


class MppiConfig(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('device', C.c_int32), ('nx', C.c_int32), ('nu', C.c_int32),
                ('num_samples', C.c_int32), ('sample_offset', C.c_int32), ('num_local', C.c_int32),
                ('max_horizon', C.c_int32), ('precision', C.c_int32), ('lambda_', C.c_double),
                ('noise_mu', C.c_double * MAX_NU), ('noise_sigma', C.c_double * (MAX_NU * MAX_NU)),
                ('has_bounds', C.c_int32), ('u_min', C.c_double * MAX_NU), ('u_max', C.c_double * MAX_NU),
                ('u_init', C.c_double * MAX_NU), ('sample_null_action', C.c_int32), ('seed', C.c_uint64),
                ('stream', C.c_void_p)]


class Partial(C.Structure):
    _fields_ = [('beta', C.c_double), ('eta', C.c_double), ('argmin', C.c_int64), ('count', C.c_int64),
                ('wsum', C.c_double * (MAX_HORIZON * MAX_NU))]


PARTIAL_DOUBLES = C.sizeof(Partial) // 8
IPC_HANDLE_BYTES = 64


def lib_path():
    return os.environ.get('TAMPC_MPPI_LIB') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc',
                                                              'libtampc_mppi.so')


_lib = None


def load():
    """Load the shared library and declare every prototype of include/tampc_mppi.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise RuntimeError("libtampc_mppi.so not found at {}; build it with `python -m tampc_b200.build` "
                           "(there is no CPU fallback)".format(path))
    lib = C.CDLL(path)
    dp, vp, i32 = C.POINTER(C.c_double), C.c_void_p, C.c_int32
    lib.tampc_mppi_last_error.restype = C.c_char_p
    lib.tampc_mppi_last_error.argtypes = [vp]
    lib.tampc_mppi_abi_version.restype = C.c_int
    lib.tampc_mppi_abi_version.argtypes = []
    lib.tampc_mppi_create.argtypes = [C.POINTER(MppiConfig), C.POINTER(vp)]
    lib.tampc_mppi_destroy.restype = None
    lib.tampc_mppi_destroy.argtypes = [vp]
    lib.tampc_mppi_set_model.argtypes = [vp, C.POINTER(ModelDesc), dp, C.c_size_t]
    lib.tampc_mppi_set_cost.argtypes = [vp, C.POINTER(CostDesc)]
    lib.tampc_mppi_set_nominal.argtypes = [vp, dp, i32]
    lib.tampc_mppi_get_nominal.argtypes = [vp, dp]
    lib.tampc_mppi_command.argtypes = [vp, dp, i32, dp, dp]
    lib.tampc_mppi_command_begin.argtypes = [vp, dp, i32, dp, i32, vp]
    lib.tampc_mppi_command_finish.argtypes = [vp, vp, i32, dp]
    lib.tampc_mppi_rollout_costs.argtypes = [vp, dp, dp, i32, dp, dp]
    lib.tampc_mppi_get_rollouts.argtypes = [vp, dp, dp]
    lib.tampc_mppi_query.argtypes = [vp, i32, dp, C.c_size_t]
    lib.tampc_mppi_select_action.argtypes = [vp, dp, dp, i32, C.POINTER(C.c_int64), dp, dp]
    lib.tampc_mppi_state_costs.argtypes = [vp, dp, i32, dp, dp]
    lib.tampc_mppi_kernel_name.restype = C.c_char_p
    lib.tampc_mppi_kernel_name.argtypes = [vp]
    lib.tampc_mppi_debug_stamps.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.tampc_mppi_device_costs.argtypes = [vp, C.POINTER(vp)]
    lib.tampc_mppi_synchronize.argtypes = [vp]
    lib.tampc_mppi_command_resident.argtypes = [vp, i32]
    lib.tampc_mppi_set_state_resident.argtypes = [vp, dp]
    lib.tampc_mppi_rollout_resident.argtypes = [vp, i32]
    lib.tampc_mppi_comm_init.argtypes = [vp, i32, i32, vp]
    lib.tampc_mppi_comm_connect.argtypes = [vp, vp]
    lib.tampc_mppi_comm_disconnect.argtypes = [vp]
    lib.tampc_mppi_predict.argtypes = [vp, dp, dp, dp]
    lib.tampc_mppi_predict_nominal.argtypes = [vp, dp, dp, dp]
    lib.tampc_mppi_set_gp.argtypes = [vp, C.POINTER(GpDesc), dp, C.c_size_t]
    lib.tampc_mppi_set_gp_draws.argtypes = [vp, dp, C.c_size_t]
    lib.tampc_gp_fit.argtypes = [i32, i32, i32, i32, dp, dp, i32, C.c_double, dp, dp, dp, C.POINTER(C.c_int64), dp]
    lib.tampc_gp_fit_last_error.restype = C.c_char_p
    lib.tampc_gp_fit_last_error.argtypes = []
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ('tampc_mppi_last_error', 'tampc_mppi_destroy', 'tampc_gp_fit_last_error', 'tampc_mppi_kernel_name'):
            fn.restype = C.c_int
    if lib.tampc_mppi_abi_version() != ABI_VERSION:
        raise RuntimeError("libtampc_mppi.so ABI {} != binding ABI {}".format(lib.tampc_mppi_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def dptr(a):
    """float64 C-contiguous numpy array -> double*."""
    assert a.dtype == np.float64 and a.flags['C_CONTIGUOUS']
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ------------------------------------------------------------------------------------------------
# spec -> descriptors

def _mask(dims):
    m = 0
    for d in dims:
        m |= 1 << int(d)
    return m


class _ParamBuilder:
    def __init__(self):
        self.chunks = []
        self.n = 0

    def add(self, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(-1)
        off = self.n
        self.chunks.append(a)
        self.n += a.size
        return off

    def mlp(self, m: S.MLPSpec):
        d = MlpDesc()
        d.n_layers = len(m.weights)
        for i, v in enumerate(m.dims):
            d.dims[i] = v
        for i, (w, b) in enumerate(zip(m.weights, m.biases)):
            d.w_off[i] = self.add(w)
            d.b_off[i] = self.add(b)
        return d

    def affine(self, a):
        d = AffineDesc()
        if a is None:
            d.scale_off = d.offset_off = -1
        else:
            d.scale_off = self.add(a.scale)
            d.offset_off = self.add(a.offset)
        return d

    def array(self):
        return np.concatenate(self.chunks) if self.chunks else np.zeros(0)


def model_desc(dyn: S.DynamicsSpec):
    """DynamicsSpec -> (tampc_model_desc, float64 params array)."""
    pb = _ParamBuilder()
    d = ModelDesc()
    d.kind, d.nx, d.nu, d.nv = dyn.kind, dyn.nx, dyn.nu, dyn.nv
    none = AffineDesc(-1, -1)
    d.aff_in = d.aff_out = d.aff_y = none
    d.ext_w_off = d.ext_b_off = -1
    slopes = set()
    if dyn.kind == S.DYN_MLP:
        d.net = pb.mlp(dyn.net)
        d.aff_in, d.aff_out = pb.affine(dyn.aff_in), pb.affine(dyn.aff_out)
        slopes.add(dyn.net.slope)
    elif dyn.kind == S.DYN_REX:
        d.net, d.encoder, d.decoder = pb.mlp(dyn.net), pb.mlp(dyn.encoder), pb.mlp(dyn.decoder)
        d.ext_w_off, d.ext_b_off, d.ne = pb.add(dyn.ext_w), pb.add(dyn.ext_b), dyn.ext_w.shape[0]
        d.aff_in, d.aff_out, d.aff_y = pb.affine(dyn.aff_in), pb.affine(dyn.aff_out), pb.affine(dyn.aff_y)
        slopes.update((dyn.net.slope, dyn.encoder.slope, dyn.decoder.slope))
    if len(slopes) > 1:
        raise ValueError("networks with different LeakyReLU slopes are not supported: {}".format(slopes))
    d.leaky_slope = slopes.pop() if slopes else 0.01
    d.wrap_mask = _mask(dyn.wrap_dims)
    d.has_bad_state_guard = int(dyn.bad_state_norm is not None)
    d.bad_state_norm = dyn.bad_state_norm if dyn.bad_state_norm is not None else 0.0
    for i, v in enumerate(dyn.pendulum):
        d.pendulum[i] = v
    return d, pb.array()


def gp_desc(gp, sampler: S.SamplerSpec):
    """GPSpec (or None = residual off) -> (tampc_gp_desc, float64 params array)."""
    d = GpDesc()
    if gp is None:
        return d, np.zeros(0)
    pb = _ParamBuilder()
    d.n_points, d.n_in, d.n_out = gp.X.shape[0], gp.X.shape[1], gp.resid.shape[1]
    if d.n_in > GP_MAX_DIM or d.n_out > GP_MAX_DIM:
        raise ValueError("GP wider than {} inputs / outputs".format(GP_MAX_DIM))
    d.space, d.sample = gp.space, int(gp.sample)
    d.x_off, d.resid_off = pb.add(gp.X), pb.add(gp.resid)
    for i in range(d.n_out):
        d.lengthscale[i], d.outputscale[i], d.noise[i] = gp.lengthscale[i], gp.outputscale[i], gp.noise[i]
    if gp.space == S.GP_ORIGINAL:
        for i in range(d.n_in):
            d.in_scale[i], d.in_offset[i] = gp.aff_in.scale[i], gp.aff_in.offset[i]
        for i in range(d.n_out):
            d.out_scale[i] = gp.aff_out.scale[i]
            if gp.aff_out.offset[i] != gp.aff_out.offset[i]:
                raise ValueError("GP output scaler offset is NaN")
    d.rollout_samples = sampler.rollout_samples
    d.rollout_var_cost, d.rollout_var_discount = sampler.rollout_var_cost, sampler.rollout_var_discount
    return d, pb.array()


def _fill(carray, values):
    """Copy a float64 numpy array into the leading entries of a (possibly nested) ctypes double array."""
    a = np.ascontiguousarray(values, dtype=np.float64)
    C.memmove(carray, a.ctypes.data, a.nbytes)


def _padded(a, shape):
    out = np.zeros(shape)
    out[tuple(slice(0, n) for n in a.shape)] = a
    return out


def cost_desc(c: S.CostSpec):
    d = CostDesc()
    d.kind, d.nx, d.nu = c.kind, c.nx, c.nu
    d.apf_max_dist = 1.0
    d.ang_mask, d.usim_mask = _mask(c.ang_dims), _mask(c.usim_dims)
    _fill(d.dist_scale, c.dist_scale)
    if c.kind == S.COST_GOAL_TRAP:
        _fill(d.goal, c.goal)
        _fill(d.Q, _padded(c.Q, (MAX_NX, MAX_NX)))
        _fill(d.R, _padded(c.R, (MAX_NU, MAX_NU)))
        d.use_trap_cost = int(c.use_trap_cost)
        d.n_traps = len(c.trap_x)
        if d.n_traps:
            _fill(d.trap_x, _padded(c.trap_x, (d.n_traps, MAX_NX)))
            _fill(d.trap_u, _padded(c.trap_u, (d.n_traps, MAX_NU)))
        d.trap_weight = c.trap_weight
        d.has_terminal = int(c.terminal_multiplier is not None)
        d.terminal_multiplier = c.terminal_multiplier if c.terminal_multiplier is not None else 0.0
    elif c.kind == S.COST_APF:
        _fill(d.goal, c.goal)
        _fill(d.Q, _padded(c.Q, (MAX_NX, MAX_NX)))
        d.n_traps = len(c.trap_x)
        if d.n_traps:
            _fill(d.trap_x, _padded(c.trap_x, (d.n_traps, MAX_NX)))
        d.apf_use_goal = int(c.apf_use_goal)
        d.apf_max_dist, d.apf_gain = float(c.apf_max_dist), float(c.apf_gain)
        if c.apf_helicoid is not None:
            rxy, oxy, cw = c.apf_helicoid
            d.apf_helicoid, d.apf_clockwise = 1, int(cw)
            d.apf_rxy[0], d.apf_rxy[1], d.apf_oxy[0], d.apf_oxy[1] = rxy[0], rxy[1], oxy[0], oxy[1]
    else:
        d.n_sets = len(c.goal_sets)
        sets = np.zeros((MAX_GOAL_SETS, MAX_SET_GOALS, MAX_NX))
        for j, (g, w) in enumerate(zip(c.goal_sets, c.set_weights)):
            d.set_size[j] = len(g)
            d.set_weight[j] = w
            sets[j, :len(g), :c.nx] = g
        _fill(d.set_goals, sets)
        d.recovery_scale = c.recovery_scale
    return d


def mppi_config(s: S.SamplerSpec, nx, precision, device=0, sample_offset=0, num_local=None, max_horizon=None,
                seed=0, stream=None):
    cfg = MppiConfig()
    cfg.abi_version = ABI_VERSION
    cfg.device = device
    cfg.nx, cfg.nu = nx, s.nu
    cfg.num_samples = s.K
    cfg.sample_offset = sample_offset
    cfg.num_local = s.K if num_local is None else num_local
    cfg.max_horizon = max(s.T, max_horizon or 0)
    cfg.precision = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
    cfg.lambda_ = s.lambda_
    for i in range(s.nu):
        cfg.noise_mu[i] = s.noise_mu[i]
        cfg.u_init[i] = s.u_init[i]
        for j in range(s.nu):
            cfg.noise_sigma[i * s.nu + j] = s.noise_sigma[i, j]
    cfg.has_bounds = int(s.u_min is not None)
    if s.u_min is not None:
        for i in range(s.nu):
            cfg.u_min[i], cfg.u_max[i] = s.u_min[i], s.u_max[i]
    cfg.sample_null_action = int(s.sample_null_action)
    cfg.seed = seed
    cfg.stream = stream
    return cfg
