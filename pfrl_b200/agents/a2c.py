"""A2C on a device-resident [T+1, num_envs] rollout window.

Public surface = pfrl/agents/a2c.py:14-310 (constructor arguments, batch
act/observe, statistics names, saved_attributes).  SURVEY section 8 (f4): the
reference already keeps the window as device tensors; what it still does per
vector step is build the reward / mask rows from Python lists and, per update,
pull three scalars to the host for the running statistics (:213-221).

Here the window is one `_Window` of preallocated tensors written in place, the
reward / done rows are uploaded as one small array each, the n-step or GAE
targets are a T-step reverse recurrence over whole [num_envs] rows (T = 5 by
default, so it stays a handful of launches), and the running statistics are
exponential averages kept ON the device and read only by get_statistics().

The draw order of the policy's sampler follows the reference (one extra
sample on the very first call, a2c.py:236-240) so that a seeded run produces
the same action sequence.

No new kernel: the fused GAE kernel (csrc/ppo.cu) is sized for PPO's
[T >= 128, E] datasets; at T = 5 the recurrence below is launch-bound either
way.
"""
import warnings

import numpy as np
import torch

from pfrl_b200 import agent
from pfrl_b200.agents.soft_actor_critic import mode_of_distribution
from pfrl_b200.utils.batch_states import batch_states
from pfrl_b200.utils.clip_l2_grad_norm import clip_l2_grad_norm_


def _row(values, device):
    """One [num_envs] fp32 row on the device from a list / array / tensor."""
    if isinstance(values, torch.Tensor):
        return values.to(device=device, dtype=torch.float32)
    return torch.from_numpy(np.asarray(values, dtype=np.float32)).to(device)


class _Window(object):
    """Rollout storage for `steps` transitions of `n_envs` environments."""

    def __init__(self, steps, n_envs, obs_shape, action_shape, device):
        f32 = dict(device=device, dtype=torch.float32)
        self.steps = steps
        self.obs_shape = tuple(obs_shape)
        self.action_shape = tuple(action_shape)
        self.states = torch.zeros((steps + 1, n_envs) + self.obs_shape, **f32)
        self.actions = torch.zeros((steps, n_envs) + self.action_shape, **f32)
        self.rewards = torch.zeros((steps, n_envs), **f32)
        self.alive = torch.ones((steps, n_envs), **f32)      # 0 where the episode ended
        self.values = torch.zeros((steps + 1, n_envs), **f32)
        self.targets = torch.zeros((steps + 1, n_envs), **f32)

    def fill_targets(self, bootstrap, gamma, gae_tau):
        """a2c.py:150-167: n-step returns, or GAE(gamma, tau) + V when gae_tau
        is given.  The carry is multiplied by `alive`, i.e. cut at episode ends."""
        T = self.steps
        if gae_tau is None:
            self.targets[T] = bootstrap
            for i in range(T - 1, -1, -1):
                self.targets[i] = self.rewards[i] + gamma * self.targets[i + 1] * self.alive[i]
            return
        self.values[T] = bootstrap
        carry = 0
        for i in range(T - 1, -1, -1):
            delta = self.rewards[i] + gamma * self.values[i + 1] * self.alive[i] - self.values[i]
            carry = delta + gamma * gae_tau * self.alive[i] * carry
            self.targets[i] = carry + self.values[i]

    def roll(self):
        self.states[0] = self.states[self.steps]


class A2C(agent.AttributeSavingMixin, agent.BatchAgent):
    """Synchronous advantage actor-critic (https://arxiv.org/abs/1708.05144).

    ``model(obs)`` must return ``(torch.distributions.Distribution, value[B, 1])``.
    """

    process_idx = None
    saved_attributes = ("model", "optimizer")

    def __init__(self, model, optimizer, gamma, num_processes, gpu=None, update_steps=5,
                 phi=lambda x: x, pi_loss_coef=1.0, v_loss_coef=0.5, entropy_coeff=0.01,
                 use_gae=False, tau=0.95, act_deterministically=False, max_grad_norm=None,
                 average_actor_loss_decay=0.999, average_entropy_decay=0.999,
                 average_value_decay=0.999, batch_states=batch_states, grad_sync=None):
        self.model = model
        if gpu is not None and gpu >= 0:
            assert torch.cuda.is_available()
            self.device = torch.device("cuda:{}".format(gpu))
            self.model.to(self.device)
        else:
            self.device = torch.device("cpu")
        self.optimizer = optimizer
        self.gamma = gamma
        self.num_processes = num_processes
        self.update_steps = update_steps
        self.phi = phi
        self.pi_loss_coef = pi_loss_coef
        self.v_loss_coef = v_loss_coef
        self.entropy_coeff = entropy_coeff
        self.use_gae = use_gae
        self.tau = tau
        self.act_deterministically = act_deterministically
        self.max_grad_norm = max_grad_norm
        self.average_actor_loss_decay = average_actor_loss_decay
        self.average_entropy_decay = average_entropy_decay
        self.average_value_decay = average_value_decay
        self.batch_states = batch_states
        self.grad_sync = grad_sync      # data-parallel hook: grad_sync(model) after backward

        self.t = 0
        self.t_start = 0
        self.window = None
        # running (actor loss, value loss, entropy) averages and their decays, on the device
        self._averages = torch.zeros(3, dtype=torch.float64, device=self.device)
        self._decays = torch.tensor(
            [average_actor_loss_decay, average_value_decay, average_entropy_decay],
            dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------- statistics
    @property
    def average_actor_loss(self):
        return float(self._averages[0])

    @property
    def average_value(self):
        return float(self._averages[1])

    @property
    def average_entropy(self):
        return float(self._averages[2])

    def get_statistics(self):
        a = self._averages.tolist()
        return [("average_actor", a[0]), ("average_value", a[1]), ("average_entropy", a[2])]

    # ------------------------------------------------------------------ update
    def _losses(self):
        w = self.window
        T, N = w.steps, self.num_processes
        pout, values = self.model(w.states[:T].reshape((T * N,) + w.obs_shape))
        actions = w.actions.reshape((T * N,) + w.action_shape)
        entropy = pout.entropy().mean()
        log_probs = pout.log_prob(actions).reshape(T, N)
        advantages = w.targets[:T] - values.reshape(T, N)
        value_loss = (advantages * advantages).mean()
        actor_loss = -(advantages.detach() * log_probs).mean()
        return actor_loss, value_loss, entropy

    def update(self):
        w = self.window
        with torch.no_grad():
            _, bootstrap = self.model(w.states[w.steps])
            w.fill_targets(bootstrap[:, 0], self.gamma, self.tau if self.use_gae else None)
        actor_loss, value_loss, entropy = self._losses()
        self.optimizer.zero_grad()
        (value_loss * self.v_loss_coef + actor_loss * self.pi_loss_coef
         - entropy * self.entropy_coeff).backward()
        if self.grad_sync is not None:
            self.grad_sync(self.model)
        if self.max_grad_norm is not None:
            clip_l2_grad_norm_(self.model.parameters(), self.max_grad_norm)
        self.optimizer.step()
        w.roll()
        self.t_start = self.t
        with torch.no_grad():
            latest = torch.stack([actor_loss, value_loss, entropy]).detach().double()
            self._averages += (1 - self._decays) * (latest - self._averages)

    # ----------------------------------------------------------------- acting
    def batch_act(self, batch_obs):
        statevar = self.batch_states(batch_obs, self.device, self.phi)
        if not self.training:
            with torch.no_grad():
                pout, _ = self.model(statevar)
                if self.act_deterministically:
                    return mode_of_distribution(pout).cpu().numpy()
                return pout.sample().cpu().numpy()
        if self.window is None:
            with torch.no_grad():
                pout, _ = self.model(statevar)
                probe = pout.sample()
            self.window = _Window(self.update_steps, self.num_processes, statevar.shape[1:],
                                  probe.shape[1:], self.device)
        w = self.window
        slot = self.t - self.t_start
        w.states[slot] = statevar
        if slot == self.update_steps:
            self.update()
            slot = 0
        with torch.no_grad():
            pout, value = self.model(statevar)
            action = pout.sample()
        w.actions[slot] = action.reshape((-1,) + w.action_shape)
        w.values[slot] = value[:, 0]
        return action.cpu().numpy()

    def batch_observe(self, batch_obs, batch_reward, batch_done, batch_reset):
        if not self.training:
            return
        self.t += 1
        ended = _row(batch_done, self.device)
        if not isinstance(batch_reset, torch.Tensor) and any(batch_reset):
            warnings.warn(
                "A2C currently does not support resetting an env without reaching a"
                " terminal state during training. When receiving True in batch_reset,"
                " A2C considers it as True in batch_done instead.")
        ended = torch.maximum(ended, _row(batch_reset, self.device))
        w = self.window
        slot = self.t - self.t_start
        w.alive[slot - 1] = 1.0 - ended
        w.rewards[slot - 1] = _row(batch_reward, self.device)
        w.states[slot] = self.batch_states(batch_obs, self.device, self.phi)
        if slot == self.update_steps:
            self.update()
