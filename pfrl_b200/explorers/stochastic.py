"""The reference's stochastic explorers: Boltzmann sampling
(pfrl/explorers/boltzmann.py) and additive Ornstein-Uhlenbeck noise
(pfrl/explorers/additive_ou.py).  Like the others they draw from numpy's
global legacy stream in the reference's order, so seeded runs agree."""
from logging import getLogger

import numpy as np
import torch

from pfrl_b200 import explorer
from pfrl_b200.explorers.epsilon_greedy import ExponentialDecayEpsilonGreedy  # NOQA (re-exported)


class Boltzmann(explorer.Explorer):
    """Sample an action with probability softmax(Q / T)."""

    def __init__(self, T=1.0):
        self.T = T

    def select_action(self, t, greedy_action_func, action_value=None):
        assert action_value is not None
        q = action_value.q_values
        with torch.no_grad():
            probs = torch.softmax(q / self.T, dim=-1).cpu().numpy().ravel()
        return np.random.choice(np.arange(q.shape[1]), p=probs)

    def __repr__(self):
        return "Boltzmann(T={})".format(self.T)


class AdditiveOU(explorer.Explorer):
    """greedy action + an Ornstein-Uhlenbeck process x += theta (mu - x) + N(0, sigma^2),
    started from its stationary law unless ``start_with_mu``."""

    def __init__(self, mu=0.0, theta=0.15, sigma=0.3, start_with_mu=False,
                 logger=getLogger(__name__)):
        self.mu = mu
        self.theta = theta
        self.sigma = sigma
        self.start_with_mu = start_with_mu
        self.logger = logger
        self.ou_state = None

    def evolve(self):
        kick = np.random.normal(size=self.ou_state.shape, loc=0, scale=self.sigma)
        self.ou_state += self.theta * (self.mu - self.ou_state) + kick

    def select_action(self, t, greedy_action_func, action_value=None):
        a = greedy_action_func()
        if self.ou_state is not None:
            self.evolve()
        elif self.start_with_mu:
            self.ou_state = np.full(a.shape, self.mu, dtype=np.float32)
        else:
            stationary = self.sigma / np.sqrt(2 * self.theta - self.theta ** 2)
            self.ou_state = np.random.normal(
                size=a.shape, loc=self.mu, scale=stationary).astype(np.float32)
        self.logger.debug("t:%s noise:%s", t, self.ou_state)
        return a + self.ou_state

    def __repr__(self):
        return "AdditiveOU(mu={}, theta={}, sigma={})".format(self.mu, self.theta, self.sigma)
