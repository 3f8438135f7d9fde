"""GPU: closed-loop behaviour, the level the reference's own scripts check by eye (scripts/pendulum.py:83-105) — the
drop-in controller, fed back its own actions through the true dynamics, swings the pendulum up and holds it; and a
push-task rollout driven by repeated commands moves the nominal-model state towards the goal."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mppi_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('precision', ['f64', 'tf32x3'])
def test_pendulum_swing_up(cuda_lib, precision):
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden('pendulum')
    spec.sampler.K = 1000
    m = B200MPPI(spec, precision=precision, seed=3)
    x = np.array([math.pi, 1.0])          # hanging down, pendulum.py:164-165
    costs = []
    for step in range(200):
        u = m.command(x).numpy()
        nxt, _ = O.apply_dynamics(spec.dynamics, torch.from_numpy(x).view(1, -1), torch.from_numpy(u).view(1, -1))
        x = nxt[0].numpy().copy()
        x[0] = (x[0] + math.pi) % (2 * math.pi) - math.pi   # the script's angle_normalize (pendulum.py:245-247)
        costs.append(x[0] ** 2 + 0.1 * x[1] ** 2)
    assert max(costs[:40]) > 5.0          # it really started at the bottom
    assert np.mean(costs[-30:]) < 0.05, np.mean(costs[-30:])   # upright and slow at the end
    m.close()


def test_push_commands_drive_the_model_towards_the_goal(cuda_lib):
    """Receding horizon on the learned model itself: 25 commands, each applied to the nominal dynamics."""
    from tampc_b200.mppi import B200MPPI
    z, spec = load_golden('push_nominal')
    spec.sampler.K = 2048
    spec.cost.trap_x, spec.cost.trap_u = np.zeros((0, 5)), np.zeros((0, 3))
    m = B200MPPI(spec, precision='tf32x3', seed=1)
    x = z['in.state'].copy()
    goal = spec.cost.goal
    d0 = np.linalg.norm(x[:2] - goal[:2])
    for _ in range(25):
        u = m.command(x)
        x = m.predict_next_state(x, u)[0]
    d1 = np.linalg.norm(x[:2] - goal[:2])
    assert d1 < 0.8 * d0, (d0, d1)
    m.close()
