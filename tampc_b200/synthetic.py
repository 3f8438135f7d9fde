"""Synthetic problems of the benchmark sweep (SURVEY.md §8d config 4 / BASELINE.json configs[3]): nx=5, nu=3, plain
residual MLP (nx+nu)->H->H->nx with LeakyReLU(0.01), QR cost + three traps."""
import numpy as np

from . import spec as S


def make_synthetic_spec(H=32, K=4096, T=40, seed=0, n_traps=3, lambda_=0.01, H2=None):
    rng = np.random.RandomState(seed)
    nx, nu = 5, 3

    def layer(out, inp, scale=1.0):
        return rng.randn(out, inp) * scale / np.sqrt(inp), rng.randn(out) * 0.01

    H2 = H if H2 is None else H2  # make_sequential_network(h_units=(H, H2)): the two hidden widths need not be equal
    w1, b1 = layer(H, nx + nu)
    w2, b2 = layer(H2, H)
    w3, b3 = layer(nx, H2, 0.1)  # x0.1 so that rollouts stay bounded
    dyn = S.DynamicsSpec(kind=S.DYN_MLP, nx=nx, nu=nu, net=S.MLPSpec([w1, w2, w3], [b1, b2, b3], slope=0.01),
                         wrap_dims=(2,), bad_state_norm=1e5)
    trng = np.random.RandomState(seed + 1)
    cost = S.CostSpec(kind=S.COST_GOAL_TRAP, nx=nx, nu=nu, ang_dims=(2,), Q=np.diag([10., 10., 0., 0., 0.]),
                      R=0.01 * np.eye(nu), goal=np.array([0.5, -0.5, 0., 0., 0.]),
                      trap_x=trng.uniform(-1, 1, (n_traps, nx)), trap_u=trng.uniform(-1, 1, (n_traps, nu)),
                      trap_weight=1.0, dist_scale=np.array([1., 1., 0.1225, 0., 0.]), usim_dims=(0, 2),
                      terminal_multiplier=50.0)
    sampler = S.SamplerSpec(K=K, T=T, nu=nu, lambda_=lambda_, noise_mu=np.array([0., 0.1, 0.]),
                            noise_sigma=np.diag([0.2, 0.4, 0.7]), u_min=np.array([-1., 0., -1.]),
                            u_max=np.array([1., 1., 1.]), u_init=np.array([0., 0.5, 0.]))
    x0 = np.random.RandomState(seed + 2).uniform(-1, 1, nx)
    return S.RolloutSpec(dyn, cost, sampler, name="synthetic_H{}".format(H) if H2 == H else "synthetic_H{}x{}".format(H, H2)), x0
