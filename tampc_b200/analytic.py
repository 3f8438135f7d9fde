"""Analytic dynamics usable as the `dynamics` of an MPC controller, carrying their own DynamicsSpec.

`Pendulum` follows scripts/pendulum.py:223-247 (`true_dynamics` + `constrain_state`): g=10, m=l=1, dt=0.05,
|u|<=2, |thdot|<=8, theta wrapped to [-pi,pi) after each step."""
import math
import torch
from . import spec as S


class Pendulum:
    def __init__(self, g=10.0, m=1.0, l=1.0, dt=0.05, u_max=2.0, thdot_max=8.0):
        self.params = (float(g), float(m), float(l), float(dt), float(u_max), float(thdot_max))

    def dynamics_spec(self) -> S.DynamicsSpec:
        return S.DynamicsSpec(kind=S.DYN_PENDULUM, nx=2, nu=1, wrap_dims=(0,), bad_state_norm=None,
                              pendulum=self.params)

    def reset(self):
        pass

    def __call__(self, state, u):
        """Host evaluation (float64 torch), used only outside the sampled rollout (e.g. MPC.__init__'s model-error
        estimate, controller.py:206-214, and predict_next_state)."""
        g, m, l, dt, umax, thdmax = self.params
        th, thdot = state[:, 0:1], state[:, 1:2]
        uc = torch.clamp(u, -umax, umax)
        newthdot = thdot + (-3 * g / (2 * l) * torch.sin(th + math.pi) + 3. / (m * l ** 2) * uc) * dt
        newth = th + newthdot * dt
        newthdot = torch.clamp(newthdot, -thdmax, thdmax)
        newth = torch.remainder(newth + math.pi, 2 * math.pi) - math.pi
        return torch.cat((newth, newthdot), dim=1)
