"""CPU oracle for the MPPI rollout hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64 torch-on-CPU restatement of what one `mpc.command(state)` computes in the reference, working from a
`tampc_b200.spec.RolloutSpec` (plain arrays) instead of live controller objects so that it can travel to the
GPU box, where /root/reference does not exist.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

torch (CPU, float64) rather than numpy because the reference itself is torch float64: the same ATen kernels
give the closest arithmetic, and bench.py's CPU baseline then times the same op mix the reference runs.

Pinned (tests/test_oracle_vs_reference.py, tests/golden/*.npz): against the reference's UNMODIFIED in-tree
classes run over oracle/shim (ExperimentalMPPI, OnlineMPC._apply_dynamics, HybridDynamicsModel,
NetworkModelWrapper, InvariantTransformer/RexExtract, cost.py, env static fns) on the committed checkpoints.
Upstream-unpinned: the pytorch_mppi softmax update and the scaler fit live in libraries that are not under
/root/reference (see oracle/shim/README.md) — "parity unpinned" for those two.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
import math

import numpy as np
import torch

from tampc_b200 import spec as S

DT = torch.float64
# Where the oracle's tensors live.  'cpu' everywhere except bench.py's secondary comparator, which re-times the same
# eager float64 op sequence on the GPU — what the reference's authors actually ran (tampc/util.py:8 get_device()).
DEVICE = 'cpu'


def _t(a):
    if torch.is_tensor(a):
        return a.to(device=DEVICE, dtype=DT)
    return torch.as_tensor(np.asarray(a), dtype=DT, device=DEVICE)


def _leaky(x, slope):
    return torch.nn.functional.leaky_relu(x, negative_slope=slope)


def mlp_forward(m: S.MLPSpec, x):
    """make_sequential_network: Linear, act, ..., Linear (model.py:72-74 user.sample)."""
    n = len(m.weights)
    for i in range(n):
        x = torch.nn.functional.linear(x, _t(m.weights[i]), _t(m.biases[i]))
        if i + 1 < n:
            x = _leaky(x, m.slope)
    return x


def angle_normalize(a):
    """arm_pytorch_utilities.math_utils.angle_normalize: ((a+pi) mod 2pi) - pi (push_main.py:192)."""
    return torch.remainder(a + math.pi, 2 * math.pi) - math.pi


def compare_to_goal(c: S.CostSpec, X, goal):
    """env.state_difference: x - goal with single-wrap angular dims (block_push.py:954-960, peg_in_hole.py:91-97,
    pendulum.py:66-71)."""
    d = X - goal
    for k in c.ang_dims:
        col = d[:, k].clone()
        col[col > math.pi] -= 2 * math.pi
        col[col < -math.pi] += 2 * math.pi
        d[:, k] = col
    return d


def state_dist(c: S.CostSpec, diff):
    """env.state_distance as a weighted 2-norm (block_push.py:663-666 r_g-scaled yaw; peg_in_hole.py:103-105)."""
    return (diff * _t(c.dist_scale)).norm(dim=1)


def u_similarity(c: S.CostSpec, U, goal_u):
    """env.control_similarity: cosine similarity on a dim subset, clamped to [0,1] (block_push.py:980-986,
    peg_in_hole.py:117-121, cost.py:69-70)."""
    dims = list(c.usim_dims)
    return torch.cosine_similarity(U[:, dims], goal_u.view(1, -1)[:, dims], dim=-1).clamp(0, 1)


# ------------------------------------------------------------------------------------------------
# dynamics

def gp_posterior(d: S.DynamicsSpec, z):
    """Exact-GP posterior of OnlineGPMixing with independent outputs at the transformed inputs z (B, n_in)
    (online_model.py:286-315 MixingBatchGP, :435-438 _make_prediction = likelihood(gp(xu)) in eval mode).
    gpytorch 1.1.1 is not under /root/reference: restated from Rasmussen & Williams eqs. 2.25-2.26 (PARITY
    UNPINNED upstream; pinned against the reference's OnlineGPMixing run over oracle/shim/gpytorch).
    Returns (mean - m(z), predictive variance incl. likelihood noise), both (B, n_out); the mean function m is
    the nominal network (online_model.py:295-302)."""
    g = d.gp
    X, resid = _t(g.X), _t(g.resid)  # resid = y - m(X), the targets minus the mean function at the training inputs
    mean = torch.zeros(z.shape[0], resid.shape[1], dtype=DT, device=z.device)
    var = torch.zeros_like(mean)
    d2_xx = torch.cdist(X, X).square()
    d2_zx = (z.unsqueeze(1) - X.unsqueeze(0)).square().sum(-1)
    for k in range(resid.shape[1]):
        l2, s, n = float(g.lengthscale[k]) ** 2, float(g.outputscale[k]), float(g.noise[k])
        Kxx = s * torch.exp(-0.5 * d2_xx / l2) + n * torch.eye(X.shape[0], dtype=DT, device=X.device)
        Kzx = s * torch.exp(-0.5 * d2_zx / l2)
        L = torch.linalg.cholesky(Kxx)
        alpha = torch.cholesky_solve(resid[:, k:k + 1], L)
        mean[:, k] = (Kzx @ alpha).squeeze(1)
        w = torch.linalg.solve_triangular(L, Kzx.T, upper=False)
        var[:, k] = s - w.square().sum(0) + n
    return mean, var


def _net_with_gp(d: S.DynamicsSpec, z, gp_eps):
    """DeterministicUser.sample (model.py:72-74), or OnlineGPMixing._dynamics_in_transformed_space
    (online_model.py:440-449): Normal(mean, stddev).sample() with the standard-normal draw `gp_eps` injected."""
    y = mlp_forward(d.net, z)
    if d.gp is None or d.gp.space != S.GP_AT_NETWORK:
        return y
    mean, var = gp_posterior(d, z)
    y = y + mean
    if d.gp.sample:
        y = y + var.sqrt() * gp_eps
    return y


def predict_dx(d: S.DynamicsSpec, x, u, gp_eps=None):
    """State difference predicted by the nominal network chain (model.py:147-165 with get_next_state=False), plus
    the GP residual when it lives in the network's space."""
    xu = torch.cat((x, u), dim=1)
    if d.kind == S.DYN_MLP:
        # PytorchTransformer(RobustMinMaxScaler) on both sides (util.py:422-423)
        z = xu * _t(d.aff_in.scale) + _t(d.aff_in.offset)
        y = _net_with_gp(d, z, gp_eps)
        return (y - _t(d.aff_out.offset)) / _t(d.aff_out.scale)
    # DYN_REX: Compose[ y-scaler, InvariantTransformer(RexExtract), z/v scaler ]
    z = mlp_forward(d.encoder, xu)  # xu_to_z, invariant.py:792-795
    z = z * _t(d.aff_in.scale) + _t(d.aff_in.offset)  # stage 3 transform_x
    v = _net_with_gp(d, z, gp_eps)  # latent dynamics, model.py:300-301 (+ GP residual)
    v = (v - _t(d.aff_out.offset)) / _t(d.aff_out.scale)  # stage 3 invert_transform
    e = torch.nn.functional.linear(x, _t(d.ext_w), _t(d.ext_b))  # x_extractor, invariant.py:885
    C = mlp_forward(d.decoder, e).view(-1, d.nx, d.nv)  # get_dx, invariant.py:926-931
    dxs = (C @ v.unsqueeze(2)).squeeze(2)  # linalg.batch_batch_product
    return (dxs - _t(d.aff_y.offset)) / _t(d.aff_y.scale)  # stage 1 invert_transform


def predict(d: S.DynamicsSpec, x, u, gp_eps=None):
    """NetworkModelWrapper.predict on cat(x,u) -> next state (model.py:147-170,131-133,300-301), with the
    preprocessor chain of util.py:408-411 and RexExtract (invariant.py:628-640,792-795,926-931).  With d.gp:
    OnlineDynamicsModel.predict (online_model.py:67-99); gp_eps (B, n_out) are the injected standard normals."""
    if d.kind == S.DYN_PENDULUM:
        # scripts/pendulum.py:223-241 true_dynamics
        g, m, l, dt, umax, thdmax = d.pendulum
        th, thdot = x[:, 0:1], x[:, 1:2]
        uc = torch.clamp(u, -umax, umax)
        newthdot = thdot + (-3 * g / (2 * l) * torch.sin(th + np.pi) + 3. / (m * l ** 2) * uc) * dt
        newth = th + newthdot * dt
        newthdot = torch.clamp(newthdot, -thdmax, thdmax)
        return torch.cat((newth, newthdot), dim=1)
    if d.gp is not None and d.gp.space == S.GP_ORIGINAL:
        # temporary local model (hybrid_model.py:189-199): the GP sits behind its own per-dim scalers and its mean
        # function is transform_y(nominal dx in the original space) (online_model.py:304-311)
        g = d.gp
        dx_nom = predict_dx(d, x, u)
        zin = torch.cat((x, u), dim=1) * _t(g.aff_in.scale) + _t(g.aff_in.offset)
        y = dx_nom * _t(g.aff_out.scale) + _t(g.aff_out.offset)
        mean, var = gp_posterior(d, zin)
        y = y + mean
        if g.sample:
            y = y + var.sqrt() * gp_eps
        return x + (y - _t(g.aff_out.offset)) / _t(g.aff_out.scale)
    return x + predict_dx(d, x, u, gp_eps)  # advance_pred_all_diff, model.py:131-133


def apply_dynamics(d: S.DynamicsSpec, state, u, gp_eps=None):
    """OnlineMPC._apply_dynamics (online_controller.py:60-79) with AlwaysSelectNominal gating.

    Returns (next_state, u_seen_by_cost): rows whose ||state|| > 1e5 have state AND action zeroed in place in the
    reference (through views), come back as zeros, and the running cost of that step sees u = 0."""
    state = state.clone()
    u = u.clone()
    if d.bad_state_norm is not None:
        bad = state.norm(dim=1) > d.bad_state_norm
        state[bad] = 0
        u[bad] = 0
    else:
        bad = None
    nxt = predict(d, state, u, gp_eps)
    for k in d.wrap_dims:  # constrain_state (push_main.py:190-193)
        nxt[:, k] = angle_normalize(nxt[:, k])
    if bad is not None:
        nxt[bad] = state[bad]  # == 0
    return nxt, u


# ------------------------------------------------------------------------------------------------
# cost

def qr_cost(c: S.CostSpec, X, U=None, terminal=False):
    """cost.qr_cost / CostQROnlineTorch (cost.py:11-36)."""
    d = compare_to_goal(c, X, _t(c.goal).view(1, -1))
    cost = torch.einsum('ij,ij->i', d @ _t(c.Q), d)
    if U is not None and not terminal:
        cost = cost + torch.einsum('ij,ij->i', U @ _t(c.R), U)
    return cost


def trap_cost(c: S.CostSpec, X, U=None, terminal=False):
    """TrapSetCost (cost.py:154-189) with MPC._trap_cost_reduce (controller.py:245-246)."""
    total = torch.zeros(X.shape[0], dtype=DT, device=X.device)
    if len(c.trap_x) == 0:
        return total
    costs = []
    for gx, gu in zip(c.trap_x, c.trap_u):
        c_x = compare_to_goal(c, X, _t(gx).view(1, -1))
        c_x = state_dist(c, c_x).square()
        c_x = torch.clamp(1 / c_x, 0, S.TRAP_MAX_COST)
        if terminal:
            c_u = 0
        elif U is None:
            c_u = 1
        else:
            c_u = u_similarity(c, U, _t(gu))
        costs.append(c_x * c_u)
    return torch.stack(costs, dim=0).sum(dim=0) * c.trap_weight


def recovery_cost(c: S.CostSpec, X):
    """GoalSetCost with min_cost inside ComposeCost(weights) (cost.py:114-147,65-66,213-235;
    online_controller.py:370-396)."""
    total = torch.zeros(X.shape[0], dtype=DT, device=X.device)
    for G, w in zip(c.goal_sets, c.set_weights):
        per_goal = []
        for g in G:
            diff = compare_to_goal(c, X, _t(g).view(1, -1))
            per_goal.append(state_dist(c, diff).square() * c.recovery_scale)
        total = total + torch.stack(per_goal, dim=0).min(dim=0).values * w
    return total


def apf_cost(c: S.CostSpec, X):
    """The APF baselines' potential of the next states (U=None): APFVO's ComposeCost([goal_cost, ArtificialRepulsionCost])
    (online_controller.py:561-564, cost.py:236-265), APFSP's goal_cost | APFHelicoid2D (online_controller.py:639-662,
    cost.py:268-289)."""
    total = torch.zeros(X.shape[0], dtype=DT, device=DEVICE)
    if c.apf_use_goal:
        d = compare_to_goal(c, X, _t(c.goal))
        total = total + torch.einsum('ni,ij,nj->n', d, _t(c.Q), d)  # qr_cost with U=None (cost.py:11-18)
    rep = torch.zeros_like(total)
    for s in c.trap_x:
        rho = state_dist(c, compare_to_goal(c, X, _t(s)))
        inside = rho <= c.apf_max_dist
        v = torch.zeros_like(rho)
        v[inside] = (1 / rho[inside] - 1 / c.apf_max_dist).square()
        rep = rep + torch.clamp(v, 0, S.TRAP_MAX_COST)
    total = total + rep * c.apf_gain
    if c.apf_helicoid is not None:
        rxy, oxy, cw = c.apf_helicoid
        orig = _t(oxy) - _t(rxy)
        orig = torch.atan2(orig[1], orig[0])
        d = _t(oxy) - X[:, :2]
        ang = orig - torch.atan2(d[:, 1], d[:, 0])
        ang = torch.where(ang > math.pi, ang - 2 * math.pi, ang)   # math_utils.angular_diff_batch: one wrap
        ang = torch.where(ang < -math.pi, ang + 2 * math.pi, ang)
        total = total + ang * torch.norm(d, dim=1) * (1 if cw else -1)
    return total


def select_action(spec: S.RolloutSpec, state, actions):
    """APFVO / APFSP action selection (online_controller.py:583-590, 657-664): argmin of the potential of the states the
    candidate actions lead to.  Returns (index, cost (n,), next_states (n, nx))."""
    x = _t(state).view(1, -1).repeat(actions.shape[0], 1)
    nxt, _ = apply_dynamics(spec.dynamics, x, _t(actions))
    c = apf_cost(spec.cost, nxt)
    return int(torch.argmin(c)), c, nxt


def normalize_trapset_cost_to_state(c: S.CostSpec, x):
    """OnlineMPPI.normalize_trapset_cost_to_state (online_controller.py:517-530): goal_cost(x) / trap_cost(x); trap_cost is
    the TrapSetCost with U=None (c_u = 1, cost.py:177) reduced by MPC._trap_cost_reduce = sum * the CURRENT trap_set_weight
    (controller.py:245-246)."""
    x = _t(x).view(1, -1)
    d = compare_to_goal(c, x, _t(c.goal))
    goal = torch.einsum('ni,ij,nj->n', d, _t(c.Q), d)
    trap = torch.zeros(1, dtype=DT, device=DEVICE)
    for tx in c.trap_x:
        cx = state_dist(c, compare_to_goal(c, x, _t(tx))).square()
        trap = trap + torch.clamp(1 / cx, 0, S.TRAP_MAX_COST)
    trap = trap * c.trap_weight
    return goal / trap, goal, trap


def running_cost(c: S.CostSpec, X, U):
    """MPC._running_cost -> ComposeCost([goal_cost, trap_cost]) (controller.py:227,248-249; cost.py:213-235)
    or OnlineMPPI._recovery_running_cost (online_controller.py:188-189)."""
    if c.kind == S.COST_RECOVERY:
        return recovery_cost(c, X)
    if c.kind == S.COST_APF:
        return apf_cost(c, X)
    cost = qr_cost(c, X, U)
    if c.use_trap_cost:
        cost = cost + trap_cost(c, X, U)
    return cost


def terminal_cost(c: S.CostSpec, X_last):
    """MPC._terminal_cost = multiplier * cost(x_T, terminal=True) (controller.py:251-259); None -> skipped
    (controller.py:376)."""
    if c.kind == S.COST_RECOVERY or c.terminal_multiplier is None:
        return torch.zeros(X_last.shape[0], dtype=DT, device=X_last.device)
    cost = qr_cost(c, X_last, None, terminal=True)
    if c.use_trap_cost:
        cost = cost + trap_cost(c, X_last, None, terminal=True)
    return c.terminal_multiplier * cost


# ------------------------------------------------------------------------------------------------
# MPPI

# True: roll a deterministic model out M = rollout_samples times too, exactly as controller.py:355-380 does (what the
# reference scripts execute at M=10); False: one replica (identical results).  bench.py's
# `reference_M10` leg sets it to time the reference's literal work.
LITERAL_REPLICAS = False


def rollout_costs_replicated(spec: S.RolloutSpec, state, perturbed_actions, gp_eps, return_states=False):
    """ExperimentalMPPI._compute_rollout_costs (controller.py:340-381) with a stochastic (GP-sampled) model: M
    replicas per sample, cost = mean over replicas + rollout_var_cost * sum_t var_M(c_t) * discount^t.
    gp_eps: (T, M, K, n_out) standard-normal draws (row m*K+k of the reference's flattened (M*K, .) batch)."""
    s = spec.sampler
    K, T, nu = perturbed_actions.shape
    M = s.rollout_samples
    x = _t(state).view(1, -1).repeat(M * K, 1)
    cost_samples = torch.zeros(M, K, dtype=DT, device=DEVICE)
    cost_var = torch.zeros(K, dtype=DT, device=DEVICE)
    states = []
    for t in range(T):
        u = perturbed_actions[:, t].repeat(M, 1)
        eps = None if gp_eps is None else _t(gp_eps[t]).reshape(M * K, -1)
        x, u_seen = apply_dynamics(spec.dynamics, x, u, eps)
        c = running_cost(spec.cost, x, u_seen).view(M, K)
        cost_samples = cost_samples + c
        if M > 1:
            cost_var = cost_var + c.var(dim=0) * (s.rollout_var_discount ** t)
        if return_states:
            states.append(x.view(M, K, -1))
    cost_samples = cost_samples + terminal_cost(spec.cost, x).view(M, K)
    total = cost_samples.mean(dim=0) + cost_var * s.rollout_var_cost
    if return_states:
        return total, torch.stack(states, dim=2)
    return total


def rollout_costs(spec: S.RolloutSpec, state, perturbed_actions, return_states=False, gp_eps=None):
    """ExperimentalMPPI._compute_rollout_costs (controller.py:340-381) for a deterministic model: all M replicas
    are identical, so mean over M = one rollout and the variance term is 0."""
    if spec.dynamics.gp is not None or (LITERAL_REPLICAS and spec.sampler.rollout_samples > 1):
        return rollout_costs_replicated(spec, state, perturbed_actions, gp_eps, return_states)
    K, T, nu = perturbed_actions.shape
    x = _t(state).view(1, -1).repeat(K, 1)
    total = torch.zeros(K, dtype=DT, device=DEVICE)
    states = []
    for t in range(T):
        x, u_seen = apply_dynamics(spec.dynamics, x, perturbed_actions[:, t])
        total = total + running_cost(spec.cost, x, u_seen)
        if return_states:
            states.append(x)
    total = total + terminal_cost(spec.cost, x)
    if return_states:
        return total, torch.stack(states, dim=1)
    return total


def perturb(spec: S.RolloutSpec, U, noise):
    """controller.py:388-395: perturbed = clamp(U + noise); post-clamp noise; lambda * noise @ Sigma^-1."""
    s = spec.sampler
    perturbed = U.view(1, *U.shape) + noise
    if s.sample_null_action:
        perturbed[s.K - 1] = 0
    if s.u_min is not None:
        perturbed = torch.max(torch.min(perturbed, _t(s.u_max)), _t(s.u_min))
    noise_post = perturbed - U
    action_cost = s.lambda_ * noise_post @ _t(s.noise_sigma_inv)
    return perturbed, noise_post, action_cost


def total_cost(spec: S.RolloutSpec, state, U, noise, gp_eps=None):
    """ExperimentalMPPI._compute_total_cost_batch (controller.py:383-402) with the noise injected."""
    perturbed, noise_post, action_cost = perturb(spec, U, noise)
    cost = rollout_costs(spec, state, perturbed, gp_eps=gp_eps)
    cost = cost + torch.sum(perturbed * action_cost, dim=(1, 2))  # NB perturbed_action, not U (controller.py:400)
    return cost, perturbed, noise_post


def shift_nominal(spec: S.RolloutSpec, U):
    """MPPI.command: U = roll(U,-1); U[-1] = u_init (pytorch_mppi)."""
    U = torch.roll(U, -1, dims=0)
    U[-1] = _t(spec.sampler.u_init)
    return U


def command(spec: S.RolloutSpec, state, U, noise, gp_eps=None):
    """One MPPI.command(state): shift, total cost, beta/eta/omega, U update; returns a dict.
    `U` is the nominal sequence BEFORE the shift; `noise` is the (K,T,nu) draw N(mu, Sigma)."""
    U = shift_nominal(spec, _t(U).clone())
    noise = _t(noise)
    cost, perturbed, noise_post = total_cost(spec, state, U, noise, gp_eps)
    beta = torch.min(cost)
    w = torch.exp(-(1.0 / spec.sampler.lambda_) * (cost - beta))
    eta = torch.sum(w)
    omega = (1. / eta) * w
    U_new = U.clone()
    for t in range(U.shape[0]):
        U_new[t] = U_new[t] + torch.sum(omega.view(-1, 1) * noise_post[:, t], dim=0)
    return {'action': U_new[0].clone(), 'U': U_new, 'cost_total': cost, 'omega': omega, 'beta': beta, 'eta': eta,
            'argmin': int(torch.argmin(cost)), 'perturbed_action': perturbed, 'noise': noise_post}


def get_rollouts(spec: S.RolloutSpec, state, U):
    """MPPI.get_rollouts(state)[0]: states after applying the current nominal U (controller.py:434-435)."""
    x = _t(state).view(1, -1)
    out = []
    for t in range(U.shape[0]):
        x, _ = apply_dynamics(spec.dynamics, x, _t(U[t]).view(1, -1))
        out.append(x[0])
    return torch.stack(out, dim=0)


def sample_noise(spec: S.RolloutSpec, seed, K=None, T=None):
    """Injected-noise convention of the parity tests: noise = mu + eps @ L^T,
    eps ~ torch.randn(K,T,nu; seed) in float64."""
    s = spec.sampler
    K = s.K if K is None else K
    T = s.T if T is None else T
    g = torch.Generator().manual_seed(seed)
    eps = _t(torch.randn(K, T, s.nu, dtype=DT, generator=g))  # drawn on the host stream, then placed on DEVICE
    return _t(s.noise_mu) + eps @ _t(s.noise_chol).T


# ------------------------------------------------------------------------------------------------
# K-sharded form of the softmax update (what the multi-GPU path computes)

def shard_partial(spec: S.RolloutSpec, cost, noise_post, global_offset=0):
    """One shard's contribution relative to ITS OWN minimum: (beta_g, eta_g, argmin_g (global index), W_g[T,nu])."""
    beta = torch.min(cost)
    w = torch.exp(-(1.0 / spec.sampler.lambda_) * (cost - beta))
    return {'beta': float(beta), 'eta': float(w.sum()), 'argmin': int(torch.argmin(cost)) + global_offset,
            'wsum': torch.einsum('k,kti->ti', w, noise_post)}


def combine_partials(spec: S.RolloutSpec, partials, U_shifted):
    """Merge shard partials in rank order with the log-sum-exp rescale s_g = exp(-(beta_g - beta)/lambda)."""
    beta = min(p['beta'] for p in partials)
    argmin = min(p['argmin'] for p in partials if p['beta'] == beta)
    scale = [math.exp(-(1.0 / spec.sampler.lambda_) * (p['beta'] - beta)) for p in partials]
    eta = sum(s * p['eta'] for s, p in zip(scale, partials))
    wsum = sum(s * p['wsum'] for s, p in zip(scale, partials))
    U_new = U_shifted + wsum / eta
    return {'beta': beta, 'eta': eta, 'argmin': argmin, 'U': U_new, 'action': U_new[0].clone()}
