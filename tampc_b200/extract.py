"""Build a RolloutSpec by introspecting a live controller.

The reference hands `pytorch_mppi.MPPI` three opaque Python callbacks (dynamics `F(state,u,t)`, running cost,
terminal cost; tampc/controller/controller.py:418-421).  A fused device kernel cannot call back into Python,
so the drop-in reads the *objects behind* those callbacks instead (SURVEY.md §8b "Callback interfaces"):

  ctrl.dynamics.nominal_model            -> NetworkModelWrapper: .model (nn.Sequential), .ds.preprocessor.tsf
  ds.preprocessor.tsf(.transforms)       -> PytorchTransformer scalers, InvariantTransformer(RexExtract)
  ctrl.goal_cost.{Q,R,goal}, ctrl.trap_set, ctrl.trap_set_weight, ctrl.terminal_cost_multiplier
  ctrl.recovery_cost                     -> GoalSetCost | ComposeCost([GoalSetCost,...], weights)
  ctrl.compare_to_goal / state_dist / trap_cost.u_similarity / constrain_state   (env static functions)

The env functions are plain callables, so they are *probed* (evaluated on basis vectors) and the inferred
closed form is then *verified* on random inputs; anything that does not match the closed forms the kernel
implements raises `UnsupportedSpec` — there is no Python/CPU fallback (BASELINE.json north_star).

Everything is duck-typed: nothing from the reference (or from oracle/) is imported here.
"""
import math
from typing import Optional

import numpy as np
import torch

from . import spec as S


class UnsupportedSpec(RuntimeError):
    """The controller uses something the fused kernel has no closed form for."""


def _np(t):
    if torch.is_tensor(t):
        return t.detach().to('cpu', torch.float64).numpy()
    return np.asarray(t, dtype=np.float64)


# ------------------------------------------------------------------------------------------------
# networks and scalers

def mlp_from_module(module) -> S.MLPSpec:
    """Flatten nn.Sequential(Linear, act, Linear, act, Sequential(Linear)) into an MLPSpec."""
    linears, slopes = [], []

    def walk(m):
        if isinstance(m, torch.nn.Linear):
            linears.append(m)
        elif isinstance(m, torch.nn.LeakyReLU):
            slopes.append(float(m.negative_slope))
        elif isinstance(m, torch.nn.ReLU):
            slopes.append(0.0)
        elif isinstance(m, (torch.nn.Sequential, torch.nn.ModuleList)):
            for c in m.children():
                walk(c)
        elif isinstance(m, (torch.nn.Identity, torch.nn.Dropout)):
            pass
        else:
            raise UnsupportedSpec("network layer {} has no fused implementation".format(type(m).__name__))

    walk(module)
    if not linears:
        raise UnsupportedSpec("no Linear layers found in {}".format(type(module).__name__))
    if len(slopes) != len(linears) - 1:
        raise UnsupportedSpec("expected an activation between every pair of Linear layers and none after the last")
    if slopes and any(s != slopes[0] for s in slopes):
        raise UnsupportedSpec("mixed activation slopes {}".format(slopes))
    for lin in linears:
        if lin.bias is None:
            raise UnsupportedSpec("Linear layers without bias are not supported")
    return S.MLPSpec([_np(l.weight) for l in linears], [_np(l.bias) for l in linears],
                     slope=slopes[0] if slopes else 0.01)


def _affine_from_single(t, n) -> Optional[S.AffineSpec]:
    """SingleTransformer -> per-dim affine (None for the identity)."""
    if t is None or type(t).__name__ == 'NullSingleTransformer':
        return None
    scale = getattr(t, '_scale', None)
    offset = getattr(t, '_min', None)
    if scale is None or offset is None:
        raise UnsupportedSpec("scaler {} is not a fitted per-dim affine map".format(type(t).__name__))
    a = S.AffineSpec(_np(scale), _np(offset))
    if a.scale.shape != (n,):
        raise UnsupportedSpec("scaler {} has {} dims, expected {}".format(type(t).__name__, a.scale.shape, n))
    # verify it really is affine
    x = torch.linspace(-1.3, 2.1, 3 * n, dtype=torch.float64).view(3, n)
    if not np.allclose(_np(t.transform(x)), _np(x) * a.scale + a.offset, rtol=1e-12, atol=1e-12):
        raise UnsupportedSpec("scaler {} is not affine".format(type(t).__name__))
    return a


def _is_pytorch_transformer(t):
    return hasattr(t, 'method_x') and hasattr(t, 'method_y')


def _is_invariant_transformer(t):
    return hasattr(t, 'tsf') and hasattr(t.tsf, 'xu_to_z') and hasattr(t.tsf, 'get_dx')


_PROBE_CACHE = {}


def _cached_probe(fn):
    """Probing costs a few host milliseconds; the env callables never change, so remember the answer per callable
    (the cache holds a reference to the callable, which keeps its id unique)."""
    import functools

    @functools.wraps(fn)
    def wrapper(f, n):
        key = (fn.__name__, id(getattr(f, '__func__', f)), id(getattr(f, '__self__', None)), n)
        hit = _PROBE_CACHE.get(key)
        if hit is not None:
            return hit[1]
        val = fn(f, n)
        _PROBE_CACHE[key] = (f, val)
        return val

    return wrapper


@_cached_probe
def probe_wrap_dims(constrain_state, nx):
    """Which state dims `constrain_state` passes through angle_normalize (push_main.py:190-193)."""
    if constrain_state is None:
        return ()
    dims = []
    for d in range(nx):
        x = torch.zeros(1, nx, dtype=torch.float64)
        x[0, d] = 4.0
        y = _np(constrain_state(x.clone()))[0]
        if abs(y[d] - (4.0 - 2 * math.pi)) < 1e-12:
            dims.append(d)
        elif abs(y[d] - 4.0) > 1e-12:
            raise UnsupportedSpec("constrain_state is not an angle wrap on dim {}".format(d))
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(64, nx, dtype=torch.float64, generator=g) - 0.5) * 30
    want = x.clone()
    for d in dims:
        want[:, d] = torch.remainder(want[:, d] + math.pi, 2 * math.pi) - math.pi
    if not torch.allclose(constrain_state(x.clone()), want, rtol=0, atol=1e-12):
        raise UnsupportedSpec("constrain_state does not match angle_normalize on dims {}".format(dims))
    return tuple(dims)


@_cached_probe
def probe_ang_dims(compare_to_goal, nx):
    """Which dims of `compare_to_goal(x, goal)` are single-wrap angular differences (block_push.py:954-960)."""
    dims = []
    zero = torch.zeros(1, nx, dtype=torch.float64)
    for d in range(nx):
        x = zero.clone()
        x[0, d] = 4.0
        diff = _np(compare_to_goal(x, zero.clone()))
        if diff.shape != (1, nx):
            raise UnsupportedSpec("compare_to_goal returns shape {}, expected (N,{})".format(diff.shape, nx))
        if abs(diff[0, d] - (4.0 - 2 * math.pi)) < 1e-12:
            dims.append(d)
        elif abs(diff[0, d] - 4.0) > 1e-12:
            raise UnsupportedSpec("compare_to_goal is neither a plain nor an angular difference on dim {}".format(d))
    g = torch.Generator().manual_seed(1)
    a = (torch.rand(64, nx, dtype=torch.float64, generator=g) - 0.5) * 6
    b = (torch.rand(1, nx, dtype=torch.float64, generator=g) - 0.5) * 6
    want = a - b
    for d in dims:
        w = want[:, d]
        w = torch.where(w > math.pi, w - 2 * math.pi, w)
        w = torch.where(w < -math.pi, w + 2 * math.pi, w)
        want[:, d] = w
    if not torch.allclose(compare_to_goal(a.clone(), b.clone()), want, rtol=0, atol=1e-12):
        raise UnsupportedSpec("compare_to_goal does not match (x - goal) with single-wrap dims {}".format(dims))
    return tuple(dims)


@_cached_probe
def probe_dist_scale(state_dist, nx):
    """`state_dist(diff)` as a weighted 2-norm: returns s with dist = ||diff * s|| (block_push.py:663-666)."""
    s = np.zeros(nx)
    for d in range(nx):
        e = torch.zeros(1, nx, dtype=torch.float64)
        e[0, d] = 1.0
        s[d] = float(_np(state_dist(e))[0])
    g = torch.Generator().manual_seed(2)
    x = torch.randn(64, nx, dtype=torch.float64, generator=g)
    want = (x * torch.from_numpy(s)).norm(dim=1)
    if not torch.allclose(state_dist(x.clone()), want, rtol=1e-12, atol=1e-14):
        raise UnsupportedSpec("state_dist is not a per-dim weighted 2-norm (weights {})".format(s))
    return s


@_cached_probe
def probe_usim_dims(u_similarity, nu):
    """`u_similarity(u1,u2)` as cosine similarity over a subset of action dims, clamped to [0,1]."""
    if u_similarity is None:
        return tuple(range(nu))  # cost.default_similarity (cost.py:69-70)
    dims = []
    for d in range(nu):
        e = torch.zeros(1, nu, dtype=torch.float64)
        e[0, d] = 1.0
        v = float(_np(u_similarity(e.clone(), e.clone())).reshape(-1)[0])
        if abs(v - 1.0) < 1e-12:
            dims.append(d)
        elif abs(v) > 1e-12:
            raise UnsupportedSpec("u_similarity is not a cosine similarity on dim {}".format(d))
    g = torch.Generator().manual_seed(3)
    a = torch.randn(64, nu, dtype=torch.float64, generator=g)
    b = torch.randn(1, nu, dtype=torch.float64, generator=g)
    want = torch.cosine_similarity(a[:, dims], b[:, dims], dim=-1).clamp(0, 1)
    if not torch.allclose(u_similarity(a.clone(), b.clone()).view(-1), want, rtol=1e-12, atol=1e-14):
        raise UnsupportedSpec("u_similarity does not match clamped cosine similarity on dims {}".format(dims))
    return tuple(dims)


def gp_from_online_model(om) -> S.GPSpec:
    """Freeze an OnlineGPMixing-like object (tampc/dynamics/online_model.py:324-457) at its current data and
    hyper-parameters.  Duck-typed on gpytorch's names: gp.covar_module = ScaleKernel(RBFKernel) batched over the
    outputs (online_model.py:299-302), likelihood = MultitaskGaussianLikelihood (task noises + global noise)."""
    if not getattr(om, 'use_independent_outputs', False):
        raise UnsupportedSpec("only the independent-output GP (OnlineAdapt.GP_KERNEL_INDEP_OUT) is fused")
    gp, lik = getattr(om, 'gp', None), getattr(om, 'likelihood', None)
    if gp is None or lik is None or getattr(om, 'xu', None) is None or len(om.xu) == 0:
        raise UnsupportedSpec("the online GP holds no data yet (the reference cannot predict in that state either)")
    if getattr(gp, 'training', False):
        raise UnsupportedSpec("the online GP is in training mode; the reference predicts only after _fit_params")
    cm = getattr(gp, 'covar_module', None)
    base = getattr(cm, 'base_kernel', None)
    if cm is None or base is None or not hasattr(cm, 'outputscale') or not hasattr(base, 'lengthscale'):
        raise UnsupportedSpec("GP kernel {} is not ScaleKernel(RBFKernel)".format(type(cm).__name__))
    if 'rbf' not in type(base).__name__.lower():
        raise UnsupportedSpec("GP base kernel {} is not an RBF kernel".format(type(base).__name__))
    ny = int(om.y.shape[1])
    ls = _np(base.lengthscale).reshape(-1)
    osc = _np(cm.outputscale).reshape(-1)
    if ls.shape != (ny,) or osc.shape != (ny,):
        raise UnsupportedSpec("GP hyper-parameters are not one (lengthscale, outputscale) per output "
                              "(ARD lengthscales are not fused)")
    noise = np.zeros(ny)
    for attr in ('task_noises', 'noise'):
        if hasattr(lik, attr):
            noise = noise + _np(getattr(lik, attr)).reshape(-1)
    if not np.all(noise > 0):
        raise UnsupportedSpec("GP likelihood {} exposes no positive noise".format(type(lik).__name__))
    # mean function at the training inputs, from the live model's own forward (online_model.py:304-315): the
    # residual targets belong to the host-side GP maintenance (fit / update), not to the per-sample path
    with torch.no_grad():
        mean_X = _np(gp.forward(om.xu).mean).reshape(len(om.xu), ny)
    return S.GPSpec(_np(om.xu), _np(om.y) - mean_X, ls, osc, noise, sample=bool(getattr(om, 'sample_dynamics', True)))


def gp_spec_from_model(online_model, nx, nu) -> S.GPSpec:
    """GPSpec of an OnlineGPMixing nominal model, including where it sits relative to its mean network."""
    inner = online_model.prior.dyn_net
    pre_gp = getattr(getattr(online_model, 'ds', None), 'preprocessor', None)
    pre_nn = getattr(getattr(inner, 'ds', None), 'preprocessor', None)
    gp = gp_from_online_model(online_model)
    if pre_gp is not pre_nn:
        # the temporary local model (hybrid_model.py:189-199): GP behind its own per-dim scalers on (x,u) / dx,
        # mean function = the whole nominal model evaluated in the original space (online_model.py:304-311)
        tsf = getattr(pre_gp, 'tsf', pre_gp)
        if tsf is None or not _is_pytorch_transformer(tsf):
            raise UnsupportedSpec("the GP's preprocessor {} is neither the nominal model's nor a per-dim "
                                  "scaler pair".format(type(tsf).__name__))
        gp.space = S.GP_ORIGINAL
        gp.aff_in = _affine_from_single(tsf.method_x, nx + nu)
        gp.aff_out = _affine_from_single(tsf.method_y, nx)
        gp.__post_init__()
    return gp


def dynamics_from_model(nominal_model, nx, nu, constrain_state=None, bad_state_norm=1e5) -> S.DynamicsSpec:
    """DynamicsSpec of a NetworkModelWrapper-like object (tampc/dynamics/model.py:210-301), or of an
    OnlineGPMixing wrapped around one (online_model.py:324-457; its mean function is prior.dyn_net)."""
    if hasattr(nominal_model, 'prior') and not hasattr(nominal_model, 'model'):
        if not hasattr(nominal_model, 'gp'):
            raise UnsupportedSpec("online model {} is not a GP residual (OnlineLinearizeMixing and the other "
                                  "variants are not fused)".format(type(nominal_model).__name__))
        inner = getattr(nominal_model.prior, 'dyn_net', None)
        if inner is None or not hasattr(inner, 'model'):
            raise UnsupportedSpec("the GP's prior exposes no .dyn_net network")
        spec = dynamics_from_model(inner, nx, nu, constrain_state, bad_state_norm)
        spec.gp = gp_spec_from_model(nominal_model, nx, nu)
        spec.__post_init__()
        return spec
    if not hasattr(nominal_model, 'model'):
        raise UnsupportedSpec("dynamics {} exposes no .model network".format(type(nominal_model).__name__))
    net = mlp_from_module(nominal_model.model)
    wrap = probe_wrap_dims(constrain_state, nx)
    ds = getattr(nominal_model, 'ds', None)
    pre = getattr(ds, 'preprocessor', None) if ds is not None else None
    tsf = getattr(pre, 'tsf', pre) if pre is not None else None
    cfg = ds.original_config() if ds is not None and hasattr(ds, 'original_config') else None
    if cfg is not None and not (getattr(cfg, 'predict_difference', True) and getattr(cfg, 'predict_all_dims', True)):
        raise UnsupportedSpec("only predict_difference & predict_all_dims models advance as x + dx (model.py:116-133)")
    common = dict(nx=nx, nu=nu, wrap_dims=wrap, bad_state_norm=bad_state_norm)
    if tsf is None:
        return S.DynamicsSpec(kind=S.DYN_MLP, net=net, **common)
    if _is_pytorch_transformer(tsf):
        return S.DynamicsSpec(kind=S.DYN_MLP, net=net, aff_in=_affine_from_single(tsf.method_x, nx + nu),
                              aff_out=_affine_from_single(tsf.method_y, nx), **common)
    transforms = getattr(tsf, 'transforms', None)
    if transforms is None:
        raise UnsupportedSpec("preprocessor {} is not recognised".format(type(tsf).__name__))
    # Compose[ PytorchTransformer(Null, y-scaler), InvariantTransformer(RexExtract), (PytorchTransformer(z,v))? ]
    if not (2 <= len(transforms) <= 3 and _is_pytorch_transformer(transforms[0])
            and _is_invariant_transformer(transforms[1])
            and (len(transforms) == 2 or _is_pytorch_transformer(transforms[2]))):
        raise UnsupportedSpec("preprocessor chain {} is not [scaler, invariant, scaler]".format(
            [type(t).__name__ for t in transforms]))
    if _affine_from_single(transforms[0].method_x, nx + nu) is not None:
        raise UnsupportedSpec("a non-identity input scaler before the invariant transform is not supported")
    aff_y = _affine_from_single(transforms[0].method_y, nx)
    inv = transforms[1].tsf
    for attr in ('encoder', 'extracted_linear_decoder', 'x_extractor'):
        if not hasattr(inv, attr):
            raise UnsupportedSpec("invariant transform {} is not RexExtract-like (missing .{})".format(
                type(inv).__name__, attr))
    encoder = mlp_from_module(inv.encoder.model)
    decoder = mlp_from_module(inv.extracted_linear_decoder.model)
    ext = inv.x_extractor
    nz, nv = encoder.dims[-1], int(inv.nv)
    aff_in = aff_out = None
    if len(transforms) == 3:
        aff_in = _affine_from_single(transforms[2].method_x, nz)
        aff_out = _affine_from_single(transforms[2].method_y, nv)
    return S.DynamicsSpec(kind=S.DYN_REX, net=net, encoder=encoder, decoder=decoder, ext_w=_np(ext.weight),
                          ext_b=_np(ext.bias), aff_in=aff_in, aff_out=aff_out, aff_y=aff_y, nv=nv, **common)


# ------------------------------------------------------------------------------------------------
# cost program

def _same_method(a, b):
    return a is not None and b is not None and getattr(a, '__func__', a) is getattr(b, '__func__', b) and \
        getattr(a, '__self__', None) is getattr(b, '__self__', None)


def _goal_set_array(goal_set, nx):
    if torch.is_tensor(goal_set):
        return _np(goal_set).reshape(-1, nx)
    if len(goal_set) == 0:
        return np.zeros((0, nx))
    return np.stack([_np(g).reshape(nx) for g in goal_set])


def cost_from_controller(ctrl, running_cost=None, terminal_state_cost='unset') -> S.CostSpec:
    """CostSpec for what `ctrl.mpc.running_cost` / `.terminal_state_cost` currently point at.

    Recognised: ctrl._running_cost (+ ctrl._terminal_cost | None) and ctrl._recovery_running_cost
    (online_controller.py:400-401,416-417).  Anything else raises UnsupportedSpec."""
    nx, nu = int(ctrl.nx), int(ctrl.nu)
    if running_cost is None:
        running_cost = ctrl.mpc.running_cost
    if terminal_state_cost == 'unset':
        terminal_state_cost = ctrl.mpc.terminal_state_cost
    ang = probe_ang_dims(ctrl.compare_to_goal, nx)
    dist_scale = probe_dist_scale(ctrl.state_dist, nx)
    if _same_method(running_cost, getattr(ctrl, '_recovery_running_cost', None)):
        rc = ctrl.recovery_cost
        if terminal_state_cost is not None:
            raise UnsupportedSpec("recovery cost with a terminal cost is not a reference configuration")
        if hasattr(rc, 'fn'):  # ComposeCost
            sets = list(rc.fn)
            w = rc.weights() if callable(rc.weights) else rc.weights
            weights = [1.0] * len(sets) if w is None else _np(w).reshape(-1).tolist()
        else:
            sets, weights = [rc], [1.0]
        goal_sets = []
        scale = None
        for gs in sets:
            if not (hasattr(gs, 'goal_set') and hasattr(gs, 'scale')) or getattr(gs, 'goal_weights', None) is not None:
                raise UnsupportedSpec("recovery cost term {} is not an unweighted GoalSetCost".format(type(gs).__name__))
            red = getattr(gs, 'reduce', None)
            if getattr(red, '__name__', '') != 'min_cost':
                raise UnsupportedSpec("GoalSetCost.reduce must be cost.min_cost")
            if scale is not None and float(gs.scale) != scale:
                raise UnsupportedSpec("goal sets with different scales")
            scale = float(gs.scale)
            goal_sets.append(_goal_set_array(gs.goal_set, nx))
        # an empty goal set contributes zeros (cost.py:144-145): drop it, its weight multiplies nothing
        keep = [(g, w) for g, w in zip(goal_sets, weights) if len(g)]
        if not keep:
            raise UnsupportedSpec("recovery cost with only empty goal sets")
        return S.CostSpec(kind=S.COST_RECOVERY, nx=nx, nu=nu, ang_dims=ang, dist_scale=dist_scale,
                          goal_sets=[g for g, _ in keep], set_weights=[w for _, w in keep], recovery_scale=scale)
    if not _same_method(running_cost, ctrl._running_cost):
        raise UnsupportedSpec("mpc.running_cost is neither ctrl._running_cost nor ctrl._recovery_running_cost")
    if terminal_state_cost is None:
        tm = None
    elif _same_method(terminal_state_cost, ctrl._terminal_cost):
        tm = float(ctrl.terminal_cost_multiplier)
    else:
        raise UnsupportedSpec("mpc.terminal_state_cost is neither ctrl._terminal_cost nor None")
    gc = ctrl.goal_cost
    if gc.goal is None:
        raise UnsupportedSpec("controller has no goal; call set_goal first")
    use_trap = getattr(ctrl, 'trap_cost', None) is not None
    usim = ()
    trap_x = trap_u = None
    if use_trap:
        usim = probe_usim_dims(ctrl.trap_cost.u_similarity, nu)
        if getattr(ctrl.trap_cost, 'goal_weights', None) is not None:
            raise UnsupportedSpec("TrapSetCost.goal_weights are not used by the reference controllers")
        if len(ctrl.trap_set):
            traps = list(ctrl.trap_set)
            if len(traps) > S.MAX_TRAPS:
                # the reference appends a trap at every _start_recovery_mode without bound (online_controller.py:353);
                # the cost program holds TAMPC_MAX_TRAPS: keep the most recent ones rather than abort mid-episode
                import warnings
                warnings.warn("trap set of {} exceeds TAMPC_MAX_TRAPS={}: the {} most recent traps are used".format(
                    len(traps), S.MAX_TRAPS, S.MAX_TRAPS), RuntimeWarning, stacklevel=2)
                traps = traps[-S.MAX_TRAPS:]
            trap_x = np.stack([_np(x).reshape(nx) for x, _ in traps])
            trap_u = np.stack([_np(u).reshape(nu) for _, u in traps])
    tw = ctrl.trap_set_weight
    tw = float(_np(tw).reshape(-1)[0]) if (torch.is_tensor(tw) or isinstance(tw, np.ndarray)) else float(tw)
    return S.CostSpec(kind=S.COST_GOAL_TRAP, nx=nx, nu=nu, ang_dims=ang, dist_scale=dist_scale, Q=_np(gc.Q), R=_np(gc.R),
                      goal=_np(gc.goal).reshape(nx), trap_x=trap_x, trap_u=trap_u, trap_weight=tw,
                      use_trap_cost=use_trap, usim_dims=usim, terminal_multiplier=tm)


def apf_cost_from_controller(ctrl, this_cost=None) -> S.CostSpec:
    """Cost program of the APF baselines' one-step action selection (online_controller.py:554-664).

    `this_cost` is the potential the controller is about to evaluate on the next states: APFVO's `ctrl.cost` =
    ComposeCost([goal_cost, ArtificialRepulsionCost]) (default when the controller has one), or APFSP's `ctrl.goal_cost` /
    a fresh APFHelicoid2D(xy, active_obstacle).  Anything else raises UnsupportedSpec."""
    nx, nu = int(ctrl.nx), int(ctrl.nu)
    if this_cost is None:
        this_cost = getattr(ctrl, 'cost', None) or ctrl.goal_cost
    ang = probe_ang_dims(ctrl.compare_to_goal, nx)
    dist_scale = probe_dist_scale(ctrl.state_dist, nx)
    gc = ctrl.goal_cost
    if gc.goal is None:
        raise UnsupportedSpec("controller has no goal; call set_goal first")
    common = dict(kind=S.COST_APF, nx=nx, nu=nu, ang_dims=ang, dist_scale=dist_scale, Q=_np(gc.Q), goal=_np(gc.goal).reshape(nx))
    terms = list(this_cost.fn) if hasattr(this_cost, 'fn') else [this_cost]
    if hasattr(this_cost, 'fn') and getattr(this_cost, 'weights', None) is not None:
        raise UnsupportedSpec("weighted ComposeCost is not an APF configuration")
    use_goal, rep, hel = False, None, None
    for t in terms:
        if t is gc:
            use_goal = True
        elif hasattr(t, 'state_set') and hasattr(t, 'max_dist_influence'):
            if t.compare_to_goal is not ctrl.compare_to_goal and not _same_method(t.compare_to_goal, ctrl.compare_to_goal):
                raise UnsupportedSpec("ArtificialRepulsionCost uses another state difference than the controller")
            rep = t
        elif hasattr(t, 'rxy') and hasattr(t, 'oxy'):
            hel = (_np(t.rxy).reshape(2), _np(t.oxy).reshape(2), bool(getattr(t, 'clockwise', False)))
        else:
            raise UnsupportedSpec("APF potential term {} has no fused implementation".format(type(t).__name__))
    trap_x = None
    max_dist, gain = 1.0, 1.0
    if rep is not None:
        states = list(rep.state_set)
        if len(states) > S.MAX_TRAPS:
            import warnings
            warnings.warn("repulsive set of {} exceeds TAMPC_MAX_TRAPS={}: the most recent ones are used".format(len(states), S.MAX_TRAPS),
                          RuntimeWarning, stacklevel=2)
            states = states[-S.MAX_TRAPS:]
        if states:
            trap_x = np.stack([_np(x).reshape(nx) for x in states])
        max_dist, gain = float(rep.max_dist_influence), float(rep.gain)
    return S.CostSpec(trap_x=trap_x, apf_use_goal=use_goal, apf_max_dist=max_dist, apf_gain=gain, apf_helicoid=hel, **common)


def sampler_from_mppi(mpc) -> S.SamplerSpec:
    """Sampler constants from a pytorch_mppi.MPPI-like object."""
    return S.SamplerSpec(K=int(mpc.K), T=int(mpc.T), nu=int(mpc.nu), lambda_=float(mpc.lambda_),
                         noise_mu=_np(mpc.noise_mu), noise_sigma=_np(mpc.noise_sigma),
                         u_min=None if mpc.u_min is None else _np(mpc.u_min),
                         u_max=None if mpc.u_max is None else _np(mpc.u_max), u_init=_np(mpc.u_init),
                         sample_null_action=bool(mpc.sample_null_action),
                         rollout_samples=int(getattr(mpc, 'M', 1)),
                         rollout_var_cost=float(getattr(mpc, 'rollout_var_cost', 0.0)),
                         rollout_var_discount=float(getattr(mpc, 'rollout_var_discount', 0.95)))


def dynamics_from_controller(ctrl) -> S.DynamicsSpec:
    """OnlineMPC._apply_dynamics as a spec: bad-state guard, AlwaysSelectNominal gating, hybrid nominal model,
    constrain_state (online_controller.py:60-79)."""
    gating = getattr(ctrl, 'gating', None)
    if gating is not None and type(gating).__name__ != 'AlwaysSelectNominal':
        raise UnsupportedSpec("only AlwaysSelectNominal gating is fused (scripts/push_main.py:857-858)")
    dyn = ctrl.dynamics
    if hasattr(dyn, 'dynamics_spec'):
        # plug-in hook for analytic dynamics (e.g. tampc_b200.analytic.Pendulum): the object states its own spec
        return dyn.dynamics_spec()
    if hasattr(dyn, 'nominal_model'):
        if getattr(dyn, 'num_local_models', lambda: 0)() and type(gating).__name__ != 'AlwaysSelectNominal':
            raise UnsupportedSpec("local models are never selected by AlwaysSelectNominal")
        model = dyn.nominal_model
        guard = 1e5  # online_controller.py:63
    else:
        model, guard = dyn, None  # plain MPC._apply_dynamics (controller.py:269-270) has no guard
    return dynamics_from_model(model, int(ctrl.nx), int(ctrl.nu), getattr(ctrl, 'constrain_state', None), guard)


def spec_from_controller(ctrl, name="") -> S.RolloutSpec:
    return S.RolloutSpec(dynamics_from_controller(ctrl), cost_from_controller(ctrl), sampler_from_mppi(ctrl.mpc),
                         name=name)
