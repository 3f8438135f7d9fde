"""pytorch_mppi.mppi.MPPI restated (source not under /root/reference; PARITY UNPINNED upstream).

Follows Williams et al. 2017 Alg. 2 and the attribute surface tampc's ExperimentalMPPI subclass reads
(controller.py:316-402) / the controllers touch."""
import torch
from torch.distributions.multivariate_normal import MultivariateNormal


def _ensure_non_zero(cost, beta, factor):
    return torch.exp(-factor * (cost - beta))


class MPPI:
    def __init__(self, dynamics, running_cost, nx, noise_sigma, num_samples=100, horizon=15, device="cpu",
                 terminal_state_cost=None, lambda_=1., noise_mu=None, u_min=None, u_max=None, u_init=None,
                 U_init=None, u_scale=1, u_per_command=1, step_dependent_dynamics=False,
                 sample_null_action=False):
        self.d = device
        self.dtype = noise_sigma.dtype
        self.K = num_samples
        self.T = horizon
        self.nx = nx
        self.nu = 1 if len(noise_sigma.shape) == 0 else noise_sigma.shape[0]
        self.lambda_ = lambda_
        if noise_mu is None:
            noise_mu = torch.zeros(self.nu, dtype=self.dtype)
        if u_init is None:
            u_init = torch.zeros_like(noise_mu)
        if self.nu == 1:
            noise_mu = noise_mu.view(-1)
            noise_sigma = noise_sigma.view(-1, 1)
        self.u_min = u_min
        self.u_max = u_max
        self.u_scale = u_scale
        self.u_per_command = u_per_command
        if self.u_max is not None and self.u_min is None:
            self.u_min = -self.u_max
        if self.u_min is not None and self.u_max is None:
            self.u_max = -self.u_min
        if self.u_min is not None:
            if not torch.is_tensor(self.u_min):
                self.u_min = torch.tensor(self.u_min, dtype=self.dtype)
            if not torch.is_tensor(self.u_max):
                self.u_max = torch.tensor(self.u_max, dtype=self.dtype)
            self.u_min = self.u_min.to(device=self.d)
            self.u_max = self.u_max.to(device=self.d)
        self.noise_mu = noise_mu.to(self.d)
        self.noise_sigma = noise_sigma.to(self.d)
        self.noise_sigma_inv = torch.inverse(self.noise_sigma)
        self.noise_dist = MultivariateNormal(self.noise_mu, covariance_matrix=self.noise_sigma)
        self.U = U_init
        self.u_init = u_init.to(self.d)
        if self.U is None:
            self.U = self.noise_dist.sample((self.T,))
        self.step_dependency = step_dependent_dynamics
        self.F = dynamics
        self.running_cost = running_cost
        self.terminal_state_cost = terminal_state_cost
        self.sample_null_action = sample_null_action
        self.state = None
        self.cost_total = None
        self.cost_total_non_zero = None
        self.omega = None
        self.states = None
        self.actions = None

    def _dynamics(self, state, u, t):
        return self.F(state, u, t) if self.step_dependency else self.F(state, u)

    def _running_cost(self, state, u):
        return self.running_cost(state, u)

    def command(self, state):
        # shift the nominal sequence one step, fill the tail with u_init
        self.U = torch.roll(self.U, -1, dims=0)
        self.U[-1] = self.u_init
        if not torch.is_tensor(state):
            state = torch.tensor(state)
        self.state = state.to(dtype=self.dtype, device=self.d)
        cost_total = self._compute_total_cost_batch()
        beta = torch.min(cost_total)
        self.cost_total_non_zero = _ensure_non_zero(cost_total, beta, 1 / self.lambda_)
        eta = torch.sum(self.cost_total_non_zero)
        self.omega = (1. / eta) * self.cost_total_non_zero
        for t in range(self.T):
            self.U[t] += torch.sum(self.omega.view(-1, 1) * self.noise[:, t], dim=0)
        action = self.U[:self.u_per_command]
        if self.u_per_command == 1:
            action = action[0]
        return action

    def reset(self):
        self.U = self.noise_dist.sample((self.T,))

    def _bound_action(self, action):
        if self.u_max is not None:
            action = torch.max(torch.min(action, self.u_max), self.u_min)
        return action

    def _compute_total_cost_batch(self):
        """Base-class version (M=1); tampc overrides it (controller.py:383-402)."""
        self.noise = self.noise_dist.sample((self.K, self.T))
        self.perturbed_action = self.U + self.noise
        if self.sample_null_action:
            self.perturbed_action[self.K - 1] = 0
        self.perturbed_action = self._bound_action(self.perturbed_action)
        self.noise = self.perturbed_action - self.U
        action_cost = self.lambda_ * self.noise @ self.noise_sigma_inv
        K, T, nu = self.perturbed_action.shape
        cost_total = torch.zeros(K, device=self.d, dtype=self.dtype)
        state = self.state.view(1, -1).repeat(K, 1)
        states, actions = [], []
        for t in range(T):
            u = self.u_scale * self.perturbed_action[:, t]
            state = self._dynamics(state, u, t)
            cost_total += self._running_cost(state, u)
            states.append(state)
            actions.append(u)
        self.actions = torch.stack(actions, dim=1)
        self.states = torch.stack(states, dim=1)
        if self.terminal_state_cost:
            cost_total += self.terminal_state_cost(self.states, self.actions)
        cost_total += torch.sum(self.U * action_cost, dim=(1, 2))
        self.cost_total = cost_total
        return cost_total

    def get_rollouts(self, state, num_rollouts=1):
        state = state.view(-1, self.nx)
        if state.size(0) == 1:
            state = state.repeat(num_rollouts, 1)
        T = self.U.shape[0]
        states = torch.zeros((num_rollouts, T + 1, self.nx), dtype=self.U.dtype, device=self.U.device)
        states[:, 0] = state
        for t in range(T):
            states[:, t + 1] = self._dynamics(states[:, t].view(num_rollouts, -1),
                                              self.u_scale * self.U[t].view(num_rollouts, -1), t)
        return states[:, 1:]
