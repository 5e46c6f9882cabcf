"""Synthetic vector environments with the benchmark configurations' shapes.

There is no emulator / physics engine in the image, and the hot path being
measured is replay + update, so these envs only produce observations of the
right shape and statistics:

  SyntheticAtariVectorEnv   uint8 84x84 frames i.i.d. uniform, stack 4 with
                            SHARED frames (LazyFrames, like VectorFrameStack),
                            18 actions, rewards in {-1, 0, 1}, geometric
                            episode length.  ``device="cuda"`` keeps frames in
                            HBM (no host round trip on the acting path);
                            ``device="cpu"`` produces numpy frames (the
                            host-buffer end-to-end path and the CPU baseline).
  SyntheticContinuousVectorEnv  float32 N(0,1) observations, continuous actions
                            (Humanoid-like 376/17, HalfCheetah-like 17/6).
"""
import numpy as np
import torch

from pfrl_b200 import env
from pfrl_b200.utils.lazy_frames import LazyFrames


class DeviceObsList(list):
    """List of per-env observations that also carries the whole batch as one
    device tensor (``.batch``), so ``batch_states`` needs no per-env work."""

    batch = None


class SyntheticAtariVectorEnv(env.VectorEnv):
    def __init__(self, num_envs, device="cuda", seed=0, n_actions=18, frame_shape=(84, 84),
                 stack=4, mean_episode_len=1000):
        self.num_envs = num_envs
        self.device = torch.device(device)
        self.on_device = self.device.type == "cuda"
        self.n_actions = n_actions
        self.frame_shape = tuple(frame_shape)
        self.stack = stack
        self.p_done = 1.0 / mean_episode_len
        self.rng = np.random.RandomState(seed)
        if self.on_device:
            self.gen = torch.Generator(device=self.device)
            self.gen.manual_seed(seed)
        self.frames = [None] * num_envs  # per env: list of `stack` frame objects
        self.batch = None

    def _new_frames(self, n):
        shape = (n, 1) + self.frame_shape
        if self.on_device:
            return torch.randint(0, 256, shape, dtype=torch.uint8, device=self.device,
                                 generator=self.gen)
        return self.rng.randint(0, 256, size=shape, dtype=np.uint8)

    def _obs(self):
        out = DeviceObsList(LazyFrames(list(f), stack_axis=0) for f in self.frames)
        if self.on_device:
            out.batch = self.batch
        return out

    def _rebuild_batch(self):
        if self.on_device:
            self.batch = torch.stack([torch.cat(f, dim=0) for f in self.frames])

    def reset(self, mask=None):
        if mask is None:
            mask = np.zeros(self.num_envs, dtype=bool)
        idx = [i for i in range(self.num_envs) if not mask[i]]
        if idx:
            new = self._new_frames(len(idx))
            for j, i in enumerate(idx):
                self.frames[i] = [new[j]] * self.stack  # first frame repeated k times
            self._rebuild_batch()
        return self._obs()

    def step(self, actions):
        new = self._new_frames(self.num_envs)
        for i in range(self.num_envs):
            self.frames[i] = self.frames[i][1:] + [new[i]]
        if self.on_device:
            self.batch = torch.cat([self.batch[:, 1:], new], dim=1)
        rewards = self.rng.randint(-1, 2, size=self.num_envs).astype(np.float64)
        dones = self.rng.rand(self.num_envs) < self.p_done
        infos = [{} for _ in range(self.num_envs)]
        return self._obs(), rewards, dones, infos

    def seed(self, seeds):
        self.rng = np.random.RandomState(seeds[0] if seeds else 0)

    def close(self):
        pass


class SyntheticContinuousVectorEnv(env.VectorEnv):
    def __init__(self, num_envs, obs_dim, act_dim, device="cuda", seed=0, mean_episode_len=1000):
        self.num_envs = num_envs
        self.obs_dim = obs_dim
        self.act_dim = act_dim
        self.device = torch.device(device)
        self.on_device = self.device.type == "cuda"
        self.p_done = 1.0 / mean_episode_len
        self.rng = np.random.RandomState(seed)
        if self.on_device:
            self.gen = torch.Generator(device=self.device)
            self.gen.manual_seed(seed)
        self.batch = None

    def _sample(self):
        if self.on_device:
            return torch.randn((self.num_envs, self.obs_dim), device=self.device,
                               generator=self.gen)
        return self.rng.randn(self.num_envs, self.obs_dim).astype(np.float32)

    def _obs(self):
        out = DeviceObsList(self.batch[i] for i in range(self.num_envs))
        if self.on_device:
            out.batch = self.batch
        return out

    def reset(self, mask=None):
        fresh = self._sample()
        if mask is None or self.batch is None:
            self.batch = fresh
        else:
            m = np.asarray(mask, dtype=bool)
            if self.on_device:
                keep = torch.as_tensor(m, device=self.device)[:, None]
                self.batch = torch.where(keep, self.batch, fresh)
            else:
                self.batch = np.where(m[:, None], self.batch, fresh)
        return self._obs()

    def step(self, actions):
        self.batch = self._sample()
        rewards = self.rng.randn(self.num_envs)
        dones = self.rng.rand(self.num_envs) < self.p_done
        return self._obs(), rewards, dones, [{} for _ in range(self.num_envs)]

    def seed(self, seeds):
        self.rng = np.random.RandomState(seeds[0] if seeds else 0)

    def close(self):
        pass
