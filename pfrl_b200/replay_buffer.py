"""Replay-buffer interface, minibatch builder and update schedule.

Mirrors pfrl/replay_buffer.py (AbstractReplayBuffer :15-114,
batch_experiences :157-212, ReplayUpdater :290-356).  ``batch_experiences``
gathers straight from HBM when it is handed the ``DeviceExperiences`` that the
device buffers' ``sample`` returns, and otherwise follows the reference's host
algorithm on lists of transition dicts.
"""
from abc import ABCMeta, abstractmethod

import torch

from pfrl_b200.replay_buffers.device_buffer import DeviceExperiences
from pfrl_b200.utils.batch_states import batch_states


class AbstractReplayBuffer(object, metaclass=ABCMeta):
    """Interface shared by replay buffers (pfrl/replay_buffer.py:15-114)."""

    @abstractmethod
    def append(self, state, action, reward, next_state=None, next_action=None,
               is_state_terminal=False, env_id=0, **kwargs):
        raise NotImplementedError

    @abstractmethod
    def sample(self, n):
        raise NotImplementedError

    @abstractmethod
    def __len__(self):
        raise NotImplementedError

    @abstractmethod
    def save(self, filename):
        raise NotImplementedError

    @abstractmethod
    def load(self, filename):
        raise NotImplementedError

    @property
    @abstractmethod
    def capacity(self):
        raise NotImplementedError

    @abstractmethod
    def stop_current_episode(self, env_id=0):
        raise NotImplementedError


def _register_device_buffers():
    from pfrl_b200.replay_buffers import device_buffer

    from pfrl_b200.replay_buffers import host

    AbstractReplayBuffer.register(device_buffer.DeviceNStepBuffer)
    AbstractReplayBuffer.register(host.HostReplayBuffer)


_register_device_buffers()


def batch_experiences(experiences, device, phi, gamma, batch_states=batch_states):
    """Vectorise k experiences of 1..n transitions each into a dict of
    tensors: state / action of the first transition, next_state of the last,
    reward = sum_i gamma^i r_i, is_state_terminal = any, discount = gamma^len
    (pfrl/replay_buffer.py:157-212)."""
    if isinstance(experiences, DeviceExperiences):
        return experiences.batch(gamma, phi, device)
    firsts = [e[0] for e in experiences]
    lasts = [e[-1] for e in experiences]
    batch = {
        "state": batch_states([t["state"] for t in firsts], device, phi),
        "action": torch.as_tensor([t["action"] for t in firsts], device=device),
        "reward": torch.as_tensor(
            [sum((gamma ** i) * e[i]["reward"] for i in range(len(e))) for e in experiences],
            dtype=torch.float32, device=device),
        "next_state": batch_states([t["next_state"] for t in lasts], device, phi),
        "is_state_terminal": torch.as_tensor(
            [any(t["is_state_terminal"] for t in e) for e in experiences],
            dtype=torch.float32, device=device),
        "discount": torch.as_tensor(
            [gamma ** len(e) for e in experiences], dtype=torch.float32, device=device),
    }
    if all(t["next_action"] is not None for t in lasts):
        batch["next_action"] = torch.as_tensor([t["next_action"] for t in lasts], device=device)
    return batch


class ReplayUpdater(object):
    """When and how often to update from replay (pfrl/replay_buffer.py:290-356)."""

    def __init__(self, replay_buffer, update_func, batchsize, episodic_update, n_times_update,
                 replay_start_size, update_interval, episodic_update_len=None):
        assert batchsize <= replay_start_size
        if episodic_update:
            raise NotImplementedError("episodic (recurrent) updates are out of scope")
        self.replay_buffer = replay_buffer
        self.update_func = update_func
        self.batchsize = batchsize
        self.episodic_update = episodic_update
        self.episodic_update_len = episodic_update_len
        self.n_times_update = n_times_update
        self.replay_start_size = replay_start_size
        self.update_interval = update_interval
        # data-parallel runs (parallel.GradSync): the agent that owns this updater; every
        # update ends in a gradient all-reduce, so all ranks must update at the SAME
        # iterations although their replay shards cross replay_start_size at different
        # ones (n-step windows depend on each rank's episode boundaries)
        self.agent = None
        self._all_ranks_ready = False

    def update_if_necessary(self, iteration):
        ready = len(self.replay_buffer) >= self.replay_start_size
        gs = getattr(self.agent, "grad_sync", None)
        if gs is None:
            if not ready or iteration % self.update_interval != 0:
                return False
        else:
            # `iteration` advances identically on every rank, so this collective is
            # entered by all of them or by none; once every shard is ready it stays so
            if iteration % self.update_interval != 0:
                return False
            if not self._all_ranks_ready:
                self._all_ranks_ready = gs.all_ready(ready)
            if not self._all_ranks_ready:
                return False
        for _ in range(self.n_times_update):
            self.update_func(self.replay_buffer.sample(self.batchsize))
        return True
