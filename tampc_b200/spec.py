"""RolloutSpec — the flattened, POD description of one MPPI problem.

This is the host-side form of what crosses the C-ABI (include/tampc_mppi.h): the dynamics model the
reference reaches through `OnlineMPC._apply_dynamics` (tampc/controller/online_controller.py:60-79 ->
dynamics/hybrid_model.py:204-221 -> dynamics/model.py:147-170 -> transform/invariant.py:628-640,
792-795,926-931), the cost program (`tampc/cost.py:11-36,114-189,213-235` wired at
controller/controller.py:216-259 and online_controller.py:370-402) and the sampler constants of
`pytorch_mppi.MPPI.__init__` (+ `ExperimentalMPPI`, controller.py:316-321).

All arrays are float64 numpy; nothing here touches torch or CUDA.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence
import numpy as np

TRAP_MAX_COST = 100000.0  # tampc/cost.py:151, controller.py:177
COS_SIM_EPS = 1e-8  # torch.cosine_similarity default eps (cost.py:70, block_push.py:986)

MAX_NX = 8
MAX_NU = 4
MAX_TRAPS = 64
MAX_GOAL_SETS = 2
MAX_SET_GOALS = 8
MAX_LAYERS = 4


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


@dataclass
class MLPSpec:
    """Sequential(Linear, act, ..., Linear): LeakyReLU(slope) between layers, none after the last
    (arm_pytorch_utilities.model.make.make_sequential_network; checkpoint keys '0','2','4.0')."""
    weights: List[np.ndarray]  # each (out, in)
    biases: List[np.ndarray]  # each (out,)
    slope: float = 0.01

    def __post_init__(self):
        self.weights = [_f64(w) for w in self.weights]
        self.biases = [_f64(b).reshape(-1) for b in self.biases]
        assert len(self.weights) == len(self.biases) and 1 <= len(self.weights) <= MAX_LAYERS
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            assert w.ndim == 2 and b.shape == (w.shape[0],), "layer {} shape mismatch".format(i)
            if i:
                assert w.shape[1] == self.weights[i - 1].shape[0]

    @property
    def dims(self):
        return [self.weights[0].shape[1]] + [w.shape[0] for w in self.weights]

    @property
    def macs(self):
        return int(sum(w.size for w in self.weights))

    def forward(self, x):
        """float64 numpy forward, (N,in)->(N,out)."""
        for i, (w, b) in enumerate(zip(self.weights, self.biases)):
            x = x @ w.T + b
            if i + 1 < len(self.weights):
                x = np.where(x >= 0, x, self.slope * x)
        return x


@dataclass
class AffineSpec:
    """Per-dim y = x*scale + offset (RobustMinMaxScaler / MinMaxScaler after fitting)."""
    scale: np.ndarray
    offset: np.ndarray

    def __post_init__(self):
        self.scale = _f64(self.scale).reshape(-1)
        self.offset = _f64(self.offset).reshape(-1)
        assert self.scale.shape == self.offset.shape

    @staticmethod
    def identity(n):
        return AffineSpec(np.ones(n), np.zeros(n))


GP_MAX_POINTS = 64  # TAMPC_GP_MAX_POINTS (OnlineGPMixing.max_data_points defaults to 50, online_model.py:327)


GP_AT_NETWORK = 0  # GP shares the nominal network's transformed space (nominal_adapt=GP_KERNEL_INDEP_OUT)
GP_ORIGINAL = 1  # GP behind its own per-dim scalers on (x,u) and dx (use_temp_local_nominal_model, hybrid_model.py:189-199)


@dataclass
class GPSpec:
    """OnlineGPMixing with independent outputs (tampc/dynamics/online_model.py:286-315,324-457), frozen at its
    current data and hyper-parameters: per output d an exact GP with kernel s_d * exp(-|a-b|^2 / (2 l_d^2)), Gaussian
    noise n_d, and the nominal model as the mean function.

    space GP_AT_NETWORK: inputs are the rows the nominal network sees (z' for RexExtract), outputs add to what it
    emits (v') — the GP was built on the nominal model's own datasource (hybrid_model.py:110-112).
    space GP_ORIGINAL: inputs are aff_in([x,u]), the mean function is aff_out(dx_nominal(x,u)), outputs are scaled
    dx — the temporary local model of hybrid_model.py:189-199, whose preprocessor is util.no_tsf_preprocessor()."""
    X: np.ndarray  # (N, n_in) training inputs  (OnlineGPMixing.xu)
    resid: np.ndarray  # (N, n_out) training targets minus the mean function at X  (OnlineGPMixing.y - m(xu))
    lengthscale: np.ndarray  # (n_out,)
    outputscale: np.ndarray  # (n_out,)
    noise: np.ndarray  # (n_out,) likelihood noise per output (task noise + global noise)
    sample: bool = True  # OnlineGPMixing.sample_dynamics: draw from N(mean, stddev) instead of returning the mean
    space: int = GP_AT_NETWORK
    aff_in: Optional['AffineSpec'] = None  # GP_ORIGINAL: scaler on [x,u]
    aff_out: Optional['AffineSpec'] = None  # GP_ORIGINAL: scaler on dx

    def __post_init__(self):
        self.X = _f64(self.X)
        self.resid = _f64(self.resid)
        assert self.X.ndim == 2 and self.resid.ndim == 2 and len(self.X) == len(self.resid)
        if not 1 <= len(self.X) <= GP_MAX_POINTS:
            raise ValueError("GP with {} data points (supported: 1..{})".format(len(self.X), GP_MAX_POINTS))
        ny = self.resid.shape[1]
        self.lengthscale = _f64(self.lengthscale).reshape(-1)
        self.outputscale = _f64(self.outputscale).reshape(-1)
        self.noise = _f64(self.noise).reshape(-1)
        assert self.lengthscale.shape == self.outputscale.shape == self.noise.shape == (ny,)
        self.sample = bool(self.sample)
        self.space = int(self.space)
        if self.space == GP_ORIGINAL:
            self.aff_in = self.aff_in or AffineSpec.identity(self.X.shape[1])
            self.aff_out = self.aff_out or AffineSpec.identity(ny)
            assert self.aff_in.scale.shape == (self.X.shape[1],) and self.aff_out.scale.shape == (ny,)


DYN_MLP = 0  # x' = x + inv_y( net( aff_in([x,u]) ) )           (UseTsf.NO_TRANSFORM / synthetic sweep)
DYN_REX = 1  # RexExtract latent dynamics (UseTsf.REX_EXTRACT)    (push / peg checkpoints)
DYN_PENDULUM = 2  # scripts/pendulum.py:223-241 analytic dynamics


@dataclass
class DynamicsSpec:
    kind: int
    nx: int
    nu: int
    # DYN_MLP: net maps aff_in([x,u]) -> y ; dx = (y - aff_out.offset)/aff_out.scale
    # DYN_REX: z = encoder([x,u]); z' = aff_in(z); v' = net(z'); v = (v' - aff_out.offset)/aff_out.scale;
    #          e = ext_w @ x + ext_b; C = decoder(e).view(nx,nv); dxs = C @ v; dx = (dxs - aff_y.offset)/aff_y.scale
    net: Optional[MLPSpec] = None
    aff_in: Optional[AffineSpec] = None
    aff_out: Optional[AffineSpec] = None
    encoder: Optional[MLPSpec] = None
    decoder: Optional[MLPSpec] = None
    ext_w: Optional[np.ndarray] = None  # (ne, nx)
    ext_b: Optional[np.ndarray] = None
    aff_y: Optional[AffineSpec] = None
    nv: int = 0
    # constrain_state: state dims passed through angle_normalize after every step (push_main.py:190-193)
    wrap_dims: Sequence[int] = ()
    # bad-state guard threshold of OnlineMPC._apply_dynamics (online_controller.py:63); None disables it
    bad_state_norm: Optional[float] = 1e5
    # DYN_PENDULUM constants (pendulum.py:227-238): g, m, l, dt, |u| clamp, |thdot| clamp
    pendulum: Sequence[float] = (10.0, 1.0, 1.0, 0.05, 2.0, 8.0)
    # online-adapted residual on top of `net`; None = nominal network only
    gp: Optional[GPSpec] = None

    def __post_init__(self):
        assert 1 <= self.nx <= MAX_NX and 1 <= self.nu <= MAX_NU
        self.wrap_dims = tuple(int(d) for d in self.wrap_dims)
        if self.kind == DYN_MLP:
            assert self.net is not None and self.net.dims[0] == self.nx + self.nu and self.net.dims[-1] == self.nx
            self.aff_in = self.aff_in or AffineSpec.identity(self.nx + self.nu)
            self.aff_out = self.aff_out or AffineSpec.identity(self.nx)
        elif self.kind == DYN_REX:
            assert self.encoder is not None and self.decoder is not None and self.net is not None
            nz = self.encoder.dims[-1]
            self.nv = self.net.dims[-1]
            assert self.encoder.dims[0] == self.nx + self.nu and self.net.dims[0] == nz
            self.ext_w = _f64(self.ext_w)
            self.ext_b = _f64(self.ext_b).reshape(-1)
            assert self.ext_w.shape == (self.decoder.dims[0], self.nx)
            assert self.decoder.dims[-1] == self.nx * self.nv
            self.aff_in = self.aff_in or AffineSpec.identity(nz)
            self.aff_out = self.aff_out or AffineSpec.identity(self.nv)
            self.aff_y = self.aff_y or AffineSpec.identity(self.nx)
        elif self.kind == DYN_PENDULUM:
            assert self.nx == 2 and self.nu == 1
        else:
            raise ValueError("unknown dynamics kind {}".format(self.kind))
        if self.gp is not None:
            assert self.kind in (DYN_MLP, DYN_REX), "a GP residual needs a network as its mean function"
            if self.gp.space == GP_AT_NETWORK:
                assert self.gp.X.shape[1] == self.net.dims[0] and self.gp.resid.shape[1] == self.net.dims[-1]
            else:
                assert self.gp.X.shape[1] == self.nx + self.nu and self.gp.resid.shape[1] == self.nx

    @property
    def macs_per_step(self):
        """Matmul MACs per sample-step as the reference executes them."""
        if self.kind == DYN_MLP:
            return self.net.macs
        if self.kind == DYN_REX:
            return self.encoder.macs + self.net.macs + self.ext_w.size + self.decoder.macs + self.nx * self.nv
        return 10


COST_GOAL_TRAP = 0  # ComposeCost([CostQROnlineTorch, TrapSetCost]) or CostQROnlineTorch alone
COST_RECOVERY = 1  # GoalSetCost / ComposeCost([GoalSetCost, GoalSetCost], weights=MAB arm)
COST_APF = 2  # APFVO: ComposeCost([goal_cost, ArtificialRepulsionCost]); APFSP: goal_cost | APFHelicoid2D (online_controller.py:554-664)


@dataclass
class CostSpec:
    kind: int
    nx: int
    nu: int
    # compare_to_goal: state dims whose difference is wrapped once into [-pi,pi] (angular_diff_batch)
    ang_dims: Sequence[int] = ()
    # --- COST_GOAL_TRAP ---
    Q: Optional[np.ndarray] = None  # (nx,nx)
    R: Optional[np.ndarray] = None  # (nu,nu)
    goal: Optional[np.ndarray] = None  # (nx,)
    trap_x: Optional[np.ndarray] = None  # (P,nx)
    trap_u: Optional[np.ndarray] = None  # (P,nu)
    trap_weight: float = 1.0  # MPC.trap_set_weight (controller.py:245-246)
    use_trap_cost: bool = True
    # state_dist(diff) = || diff * dist_scale ||_2 (push: (1,1,r_g,0,0) block_push.py:663-666; peg (1,1,0,0,0))
    dist_scale: Optional[np.ndarray] = None
    # u_similarity: cosine similarity on these action dims clamped to [0,1] (block_push.py:980-986)
    usim_dims: Sequence[int] = ()
    # terminal: terminal_multiplier * cost(x_T, terminal=True); None <=> mpc.terminal_state_cost is None
    terminal_multiplier: Optional[float] = None
    # --- COST_RECOVERY ---
    goal_sets: Sequence[np.ndarray] = ()  # each (n_j, nx)
    set_weights: Sequence[float] = ()
    recovery_scale: float = 1.0
    # --- COST_APF (potentials of the NEXT state, U=None): Q / goal as above, repulsive states in trap_x ---
    apf_use_goal: bool = True
    apf_max_dist: float = 1.0   # ArtificialRepulsionCost.max_dist_influence (cost.py:239-243)
    apf_gain: float = 1.0
    apf_helicoid: Optional[tuple] = None  # (rxy (2,), oxy (2,), clockwise) of APFHelicoid2D (cost.py:268-289)

    def __post_init__(self):
        self.ang_dims = tuple(int(d) for d in self.ang_dims)
        self.usim_dims = tuple(int(d) for d in self.usim_dims)
        if self.dist_scale is None:
            ds = np.zeros(self.nx)
            ds[:2] = 1.0  # controller.default_state_dist (controller.py:184-185)
            self.dist_scale = ds
        self.dist_scale = _f64(self.dist_scale).reshape(-1)
        assert self.dist_scale.shape == (self.nx,)
        if self.kind == COST_GOAL_TRAP:
            self.Q = _f64(self.Q)
            self.R = _f64(self.R)
            self.goal = _f64(self.goal).reshape(-1)
            assert self.Q.shape == (self.nx, self.nx) and self.R.shape == (self.nu, self.nu)
            assert self.goal.shape == (self.nx,)
            if self.trap_x is None or len(self.trap_x) == 0:
                self.trap_x = np.zeros((0, self.nx))
                self.trap_u = np.zeros((0, self.nu))
            self.trap_x = _f64(self.trap_x).reshape(-1, self.nx)
            self.trap_u = _f64(self.trap_u).reshape(-1, self.nu)
            assert len(self.trap_x) == len(self.trap_u)
            if len(self.trap_x) > MAX_TRAPS:
                raise ValueError("trap set of {} exceeds TAMPC_MAX_TRAPS={}".format(len(self.trap_x), MAX_TRAPS))
            self.trap_weight = float(self.trap_weight)
        elif self.kind == COST_APF:
            self.Q = _f64(self.Q)
            self.goal = _f64(self.goal).reshape(-1)
            assert self.Q.shape == (self.nx, self.nx) and self.goal.shape == (self.nx,)
            self.R = np.zeros((self.nu, self.nu))
            self.trap_x = np.zeros((0, self.nx)) if self.trap_x is None or len(self.trap_x) == 0 else _f64(self.trap_x).reshape(-1, self.nx)
            self.trap_u = np.zeros((len(self.trap_x), self.nu))
            if len(self.trap_x) > MAX_TRAPS:
                raise ValueError("repulsive set of {} exceeds TAMPC_MAX_TRAPS={}".format(len(self.trap_x), MAX_TRAPS))
            if self.apf_helicoid is not None:
                rxy, oxy, cw = self.apf_helicoid
                self.apf_helicoid = (_f64(rxy).reshape(2), _f64(oxy).reshape(2), bool(cw))
        elif self.kind == COST_RECOVERY:
            self.goal_sets = [_f64(g).reshape(-1, self.nx) for g in self.goal_sets]
            self.set_weights = [float(w) for w in self.set_weights]
            assert 1 <= len(self.goal_sets) <= MAX_GOAL_SETS and len(self.set_weights) == len(self.goal_sets)
            for g in self.goal_sets:
                if len(g) > MAX_SET_GOALS:
                    raise ValueError("goal set of {} exceeds TAMPC_MAX_SET_GOALS={}".format(len(g), MAX_SET_GOALS))
        else:
            raise ValueError("unknown cost kind {}".format(self.kind))


@dataclass
class SamplerSpec:
    """pytorch_mppi.MPPI constructor constants + ExperimentalMPPI extras (controller.py:316-321)."""
    K: int
    T: int
    nu: int
    lambda_: float
    noise_mu: np.ndarray
    noise_sigma: np.ndarray  # covariance (nu,nu)
    u_min: Optional[np.ndarray] = None
    u_max: Optional[np.ndarray] = None
    u_init: Optional[np.ndarray] = None
    sample_null_action: bool = False
    rollout_samples: int = 1  # M; replicas of a deterministic model are identical
    rollout_var_cost: float = 0.0
    rollout_var_discount: float = 0.95

    def __post_init__(self):
        self.noise_mu = _f64(self.noise_mu).reshape(-1)
        self.noise_sigma = _f64(self.noise_sigma).reshape(self.nu, self.nu)
        self.u_init = _f64(self.u_init if self.u_init is not None else np.zeros(self.nu)).reshape(-1)
        if (self.u_min is None) != (self.u_max is None):  # math_utils.get_bounds mirroring
            if self.u_min is None:
                self.u_min = -_f64(self.u_max)
            else:
                self.u_max = -_f64(self.u_min)
        if self.u_min is not None:
            self.u_min = np.broadcast_to(_f64(self.u_min).reshape(-1), (self.nu,)).copy()
            self.u_max = np.broadcast_to(_f64(self.u_max).reshape(-1), (self.nu,)).copy()

    @property
    def noise_sigma_inv(self):
        return np.linalg.inv(self.noise_sigma)

    @property
    def noise_chol(self):
        """Lower-triangular L with Sigma = L L^T (MultivariateNormal's scale_tril)."""
        return np.linalg.cholesky(self.noise_sigma)


@dataclass
class RolloutSpec:
    dynamics: DynamicsSpec
    cost: CostSpec
    sampler: SamplerSpec
    name: str = ""

    def __post_init__(self):
        assert self.dynamics.nx == self.cost.nx and self.dynamics.nu == self.cost.nu == self.sampler.nu


# ------------------------------------------------------------------------------------------------
# flat (npz-able) form, used for the committed golden fixtures and for packaged specs

def _put_mlp(d, prefix, m: Optional[MLPSpec]):
    if m is None:
        return
    d[prefix + '.n'] = np.array(len(m.weights))
    d[prefix + '.slope'] = np.array(m.slope)
    for i, (w, b) in enumerate(zip(m.weights, m.biases)):
        d['{}.w{}'.format(prefix, i)] = w
        d['{}.b{}'.format(prefix, i)] = b


def _get_mlp(d, prefix):
    if prefix + '.n' not in d:
        return None
    n = int(d[prefix + '.n'])
    return MLPSpec([d['{}.w{}'.format(prefix, i)] for i in range(n)],
                   [d['{}.b{}'.format(prefix, i)] for i in range(n)], float(d[prefix + '.slope']))


def _put_aff(d, prefix, a: Optional[AffineSpec]):
    if a is not None:
        d[prefix + '.scale'] = a.scale
        d[prefix + '.offset'] = a.offset


def _get_aff(d, prefix):
    if prefix + '.scale' not in d:
        return None
    return AffineSpec(d[prefix + '.scale'], d[prefix + '.offset'])


def spec_to_arrays(spec: RolloutSpec, prefix='spec.'):
    d = {}
    dy, c, s = spec.dynamics, spec.cost, spec.sampler
    d['dyn.kind'] = np.array(dy.kind)
    d['dyn.nx'] = np.array(dy.nx)
    d['dyn.nu'] = np.array(dy.nu)
    d['dyn.wrap_dims'] = np.array(dy.wrap_dims, dtype=np.int64)
    d['dyn.bad_state_norm'] = np.array(np.nan if dy.bad_state_norm is None else dy.bad_state_norm)
    d['dyn.pendulum'] = _f64(dy.pendulum)
    _put_mlp(d, 'dyn.net', dy.net)
    _put_mlp(d, 'dyn.encoder', dy.encoder)
    _put_mlp(d, 'dyn.decoder', dy.decoder)
    _put_aff(d, 'dyn.aff_in', dy.aff_in)
    _put_aff(d, 'dyn.aff_out', dy.aff_out)
    _put_aff(d, 'dyn.aff_y', dy.aff_y)
    if dy.ext_w is not None:
        d['dyn.ext_w'] = dy.ext_w
        d['dyn.ext_b'] = dy.ext_b
    if dy.gp is not None:
        g = dy.gp
        d['dyn.gp.X'], d['dyn.gp.resid'] = g.X, g.resid
        d['dyn.gp.lengthscale'], d['dyn.gp.outputscale'], d['dyn.gp.noise'] = g.lengthscale, g.outputscale, g.noise
        d['dyn.gp.sample'], d['dyn.gp.space'] = np.array(int(g.sample)), np.array(g.space)
        _put_aff(d, 'dyn.gp.aff_in', g.aff_in)
        _put_aff(d, 'dyn.gp.aff_out', g.aff_out)
    d['cost.kind'] = np.array(c.kind)
    d['cost.ang_dims'] = np.array(c.ang_dims, dtype=np.int64)
    d['cost.dist_scale'] = c.dist_scale
    d['cost.usim_dims'] = np.array(c.usim_dims, dtype=np.int64)
    if c.kind == COST_GOAL_TRAP:
        d['cost.Q'], d['cost.R'], d['cost.goal'] = c.Q, c.R, c.goal
        d['cost.trap_x'], d['cost.trap_u'] = c.trap_x, c.trap_u
        d['cost.trap_weight'] = np.array(c.trap_weight)
        d['cost.use_trap_cost'] = np.array(int(c.use_trap_cost))
        d['cost.terminal_multiplier'] = np.array(np.nan if c.terminal_multiplier is None else c.terminal_multiplier)
    else:
        d['cost.n_sets'] = np.array(len(c.goal_sets))
        for j, g in enumerate(c.goal_sets):
            d['cost.set{}'.format(j)] = g
        d['cost.set_weights'] = _f64(c.set_weights)
        d['cost.recovery_scale'] = np.array(c.recovery_scale)
    d['mppi.K'], d['mppi.T'] = np.array(s.K), np.array(s.T)
    d['mppi.lambda'] = np.array(s.lambda_)
    d['mppi.noise_mu'], d['mppi.noise_sigma'], d['mppi.u_init'] = s.noise_mu, s.noise_sigma, s.u_init
    if s.u_min is not None:
        d['mppi.u_min'], d['mppi.u_max'] = s.u_min, s.u_max
    d['mppi.sample_null_action'] = np.array(int(s.sample_null_action))
    d['mppi.rollout_samples'] = np.array(s.rollout_samples)
    d['mppi.rollout_var_cost'] = np.array(s.rollout_var_cost)
    d['mppi.rollout_var_discount'] = np.array(s.rollout_var_discount)
    d['name'] = np.array(spec.name)
    return {prefix + k: v for k, v in d.items()}


def spec_from_arrays(arrays, prefix='spec.'):
    d = {k[len(prefix):]: arrays[k] for k in arrays.keys() if k.startswith(prefix)}
    nx, nu = int(d['dyn.nx']), int(d['dyn.nu'])
    bsn = float(d['dyn.bad_state_norm'])
    gp = None
    if 'dyn.gp.X' in d:
        gp = GPSpec(d['dyn.gp.X'], d['dyn.gp.resid'], d['dyn.gp.lengthscale'], d['dyn.gp.outputscale'],
                    d['dyn.gp.noise'], bool(int(d['dyn.gp.sample'])), int(d['dyn.gp.space']),
                    _get_aff(d, 'dyn.gp.aff_in'), _get_aff(d, 'dyn.gp.aff_out'))
    dyn = DynamicsSpec(kind=int(d['dyn.kind']), nx=nx, nu=nu, net=_get_mlp(d, 'dyn.net'), gp=gp,
                       aff_in=_get_aff(d, 'dyn.aff_in'), aff_out=_get_aff(d, 'dyn.aff_out'),
                       encoder=_get_mlp(d, 'dyn.encoder'), decoder=_get_mlp(d, 'dyn.decoder'),
                       ext_w=d.get('dyn.ext_w'), ext_b=d.get('dyn.ext_b'), aff_y=_get_aff(d, 'dyn.aff_y'),
                       wrap_dims=tuple(d['dyn.wrap_dims'].tolist()),
                       bad_state_norm=None if np.isnan(bsn) else bsn, pendulum=tuple(d['dyn.pendulum'].tolist()))
    kind = int(d['cost.kind'])
    common = dict(kind=kind, nx=nx, nu=nu, ang_dims=tuple(d['cost.ang_dims'].tolist()),
                  dist_scale=d['cost.dist_scale'], usim_dims=tuple(d['cost.usim_dims'].tolist()))
    if kind == COST_GOAL_TRAP:
        tm = float(d['cost.terminal_multiplier'])
        cost = CostSpec(Q=d['cost.Q'], R=d['cost.R'], goal=d['cost.goal'], trap_x=d['cost.trap_x'],
                        trap_u=d['cost.trap_u'], trap_weight=float(d['cost.trap_weight']),
                        use_trap_cost=bool(int(d['cost.use_trap_cost'])),
                        terminal_multiplier=None if np.isnan(tm) else tm, **common)
    else:
        n = int(d['cost.n_sets'])
        cost = CostSpec(goal_sets=[d['cost.set{}'.format(j)] for j in range(n)],
                        set_weights=d['cost.set_weights'].tolist(), recovery_scale=float(d['cost.recovery_scale']),
                        **common)
    sampler = SamplerSpec(K=int(d['mppi.K']), T=int(d['mppi.T']), nu=nu, lambda_=float(d['mppi.lambda']),
                          noise_mu=d['mppi.noise_mu'], noise_sigma=d['mppi.noise_sigma'],
                          u_min=d.get('mppi.u_min'), u_max=d.get('mppi.u_max'), u_init=d['mppi.u_init'],
                          sample_null_action=bool(int(d['mppi.sample_null_action'])),
                          rollout_samples=int(d['mppi.rollout_samples']),
                          rollout_var_cost=float(d['mppi.rollout_var_cost']),
                          rollout_var_discount=float(d['mppi.rollout_var_discount']))
    return RolloutSpec(dyn, cost, sampler, name=str(d['name']))


# ------------------------------------------------------------------------------------------------
# packaged specs: the extracted problem — network weights, FITTED scaler vectors, env tags, cost
# program, sampler constants — in one .npz beside the reference's checkpoint, so that a controller can start without
# re-fitting the RobustMinMaxScalers from the .mat datasets (tampc/util.py:408-415) and without the un-vendored
# preprocess library on the runtime path.  The reference's own checkpoint format is untouched.

PACKAGE_FORMAT = 1


def save_spec(path, spec: RolloutSpec):
    arrays = spec_to_arrays(spec)
    arrays['package.format'] = np.array(PACKAGE_FORMAT)
    np.savez_compressed(path, **arrays)


def load_spec(path) -> RolloutSpec:
    with np.load(path) as z:
        if 'package.format' not in z.files:
            raise ValueError("{} is not a packaged RolloutSpec".format(path))
        if int(z['package.format']) != PACKAGE_FORMAT:
            raise ValueError("packaged spec format {} != {}".format(int(z['package.format']), PACKAGE_FORMAT))
        return spec_from_arrays({k: z[k] for k in z.files})
