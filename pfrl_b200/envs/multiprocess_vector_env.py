"""VectorEnv with one subprocess per environment and a shared observation slab.

Public surface = pfrl/envs/multiprocess_vector_env.py:40-146 (constructor
from ``env_fns``, ``step / reset(mask) / seed / close / num_envs / spec``,
``action_space`` / ``observation_space`` taken from the first env).

SURVEY section 8 (f2): in the reference every observation travels
worker -> pickle -> pipe -> unpickle -> Python list -> ``batch_states`` ->
one pageable H2D copy.  Here, once the observation shape is known (first
``reset``), the parent allocates ONE shared-memory slab ``[num_envs, *obs]``;
workers write their observation straight into their row and only
``(reward, done, info)`` go through the pipe.  The slab can be page-locked
(``pin=True``, cudaHostRegister) so that the acting path uploads all
environments with a single DMA: the returned observation list carries the slab
as ``.host_batch`` and ``pfrl_b200.utils.batch_states`` uses it when ``phi``
declares its on-device form.  Per-environment entries of the list are private
copies (they end up inside replay transitions and must outlive the slab row).

Observations that are not fixed-shape arrays simply keep using the pipe.
"""
import mmap
import multiprocessing
import signal
import traceback
from multiprocessing import shared_memory

import numpy as np
import torch

from pfrl_b200 import env

_IN_SLAB = "__in_slab__"


class HostObsList(list):
    """Per-env observations + the whole batch as one host tensor
    (``.host_batch``, valid until the next ``step`` / ``reset``)."""

    host_batch = None


_WORKER_ERROR = "__worker_error__"


def _worker(remote, env_fn):
    signal.signal(signal.SIGINT, signal.SIG_IGN)  # the parent handles CTRL+C
    try:
        environment = env_fn()
    except BaseException:
        remote.send((_WORKER_ERROR, traceback.format_exc()))
        remote.close()
        return
    shm = slab = row = None

    def ship(ob):
        if row is not None:
            arr = np.asarray(ob)
            if arr.shape == row.shape and arr.dtype == row.dtype:
                row[...] = arr
                return _IN_SLAB
        return ob

    try:
        while True:
            cmd, data = remote.recv()
            if cmd == "step":
                ob, reward, done, info = environment.step(data)
                remote.send((ship(ob), reward, done, info))
            elif cmd == "reset":
                remote.send(ship(environment.reset()))
            elif cmd == "attach":
                name, shape, dtype, index = data
                # map the parent's POSIX segment directly: SharedMemory(name=...) would
                # register it with a resource tracker that does not own it (Python < 3.13)
                with open("/dev/shm/" + name.lstrip("/"), "r+b") as f:
                    shm = mmap.mmap(f.fileno(), 0)
                slab = np.ndarray(shape, dtype=np.dtype(dtype), buffer=shm)
                row = slab[index]
                remote.send(True)
            elif cmd == "get_spaces":
                remote.send((getattr(environment, "action_space", None),
                             getattr(environment, "observation_space", None)))
            elif cmd == "spec":
                remote.send(getattr(environment, "spec", None))
            elif cmd == "seed":
                remote.send(environment.seed(data))
            elif cmd == "close":
                remote.close()
                break
            else:
                raise NotImplementedError(cmd)
    except (EOFError, KeyboardInterrupt):
        pass  # the parent went away
    except BaseException:
        remote.send((_WORKER_ERROR, traceback.format_exc()))
    finally:
        row = slab = None
        if shm is not None:
            try:
                shm.close()
            except BufferError:
                pass
        environment.close()


class MultiprocessVectorEnv(env.VectorEnv):
    """Args:
        env_fns (list of callable): each returns the env run in its own subprocess.
        shared_obs (bool): move observations through a shared-memory slab.
        pin (bool): page-lock the slab for one-DMA uploads (needs CUDA).
        context (str or None): multiprocessing start method (default: platform default).
    """

    def __init__(self, env_fns, shared_obs=True, pin=False, context=None):
        ctx = multiprocessing.get_context(context)
        pipes = [ctx.Pipe() for _ in env_fns]
        self.remotes = [p[0] for p in pipes]
        self.work_remotes = [p[1] for p in pipes]
        self.ps = [ctx.Process(target=_worker, args=(w, fn), daemon=True)
                   for w, fn in zip(self.work_remotes, env_fns)]
        for p in self.ps:
            p.start()
        self.closed = False
        self.last_obs = [None] * self.num_envs
        self._shared_obs = shared_obs
        self._pin = pin
        self._shm = None
        self._slab = None
        self._slab_tensor = None
        self._pinned = False
        self._row_fresh = [False] * len(env_fns)   # slab row i holds env i's current observation
        self._spec = None
        self.remotes[0].send(("get_spaces", None))
        self.action_space, self.observation_space = self._recv(0)

    def __del__(self):
        if not getattr(self, "closed", True):
            self.close()

    def _recv(self, i):
        """Answer of worker i; a worker that died or raised becomes a RuntimeError
        here instead of a parent blocked on the pipe forever."""
        remote = self.remotes[i]
        while not remote.poll(0.5):
            if not self.ps[i].is_alive() and not remote.poll(0):
                raise RuntimeError("environment worker {} exited unexpectedly (exit code {})"
                                   .format(i, self.ps[i].exitcode))
        answer = remote.recv()
        if isinstance(answer, tuple) and len(answer) == 2 and answer[0] == _WORKER_ERROR:
            raise RuntimeError("environment worker {} failed:\n{}".format(i, answer[1]))
        return answer

    # ------------------------------------------------------------------ slab
    def _maybe_create_slab(self, sample):
        if not self._shared_obs or self._slab is not None:
            return
        self._shared_obs = False  # one attempt only
        if not isinstance(sample, np.ndarray) or sample.dtype == object or sample.size == 0:
            return
        shape = (self.num_envs,) + sample.shape
        self._shm = shared_memory.SharedMemory(create=True, size=int(np.prod(shape)) * sample.itemsize)
        self._slab = np.ndarray(shape, dtype=sample.dtype, buffer=self._shm.buf)
        for i, o in enumerate(self.last_obs):
            if isinstance(o, np.ndarray) and o.shape == sample.shape and o.dtype == sample.dtype:
                self._slab[i] = o
                self._row_fresh[i] = True
        for i, remote in enumerate(self.remotes):
            remote.send(("attach", (self._shm.name, shape, sample.dtype.str, i)))
        for i in range(self.num_envs):
            assert self._recv(i) is True
        self._slab_tensor = torch.from_numpy(self._slab)
        if self._pin and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(self._slab_tensor.data_ptr(),
                                                      self._slab.nbytes, 0)
            self._pinned = int(rc) == 0
        self._shared_obs = True

    def _take(self, i, shipped):
        if isinstance(shipped, str) and shipped == _IN_SLAB:
            self._row_fresh[i] = True
            return self._slab[i].copy()
        self._row_fresh[i] = False
        return shipped

    def _wrap(self, obs):
        out = HostObsList(obs)
        if self._slab is not None and all(self._row_fresh):
            out.host_batch = self._slab_tensor
        return out

    # ------------------------------------------------------------- VectorEnv
    @property
    def spec(self):
        if self._spec is None:
            self._assert_not_closed()
            self.remotes[0].send(("spec", None))
            self._spec = self._recv(0)
        return self._spec

    def step(self, actions):
        self._assert_not_closed()
        for remote, action in zip(self.remotes, actions):
            remote.send(("step", action))
        results = [self._recv(i) for i in range(self.num_envs)]
        obs, rews, dones, infos = zip(*results)
        self.last_obs = [self._take(i, o) for i, o in enumerate(obs)]
        return self._wrap(self.last_obs), rews, dones, infos

    def reset(self, mask=None):
        self._assert_not_closed()
        if mask is None:
            mask = np.zeros(self.num_envs)
        for keep, remote in zip(mask, self.remotes):
            if not keep:
                remote.send(("reset", None))
        self.last_obs = [o if keep else self._take(i, self._recv(i))
                         for i, (keep, o) in enumerate(zip(mask, self.last_obs))]
        first = next((o for o in self.last_obs if o is not None), None)
        self._maybe_create_slab(first)
        return self._wrap(self.last_obs)

    def seed(self, seeds=None):
        self._assert_not_closed()
        if seeds is None:
            seeds = [None] * self.num_envs
        elif isinstance(seeds, int):
            seeds = [seeds] * self.num_envs
        elif isinstance(seeds, list):
            if len(seeds) != self.num_envs:
                raise ValueError(
                    "length of seeds must be same as num_envs {}".format(self.num_envs))
        else:
            raise TypeError("Type of Seeds {} is not supported.".format(type(seeds)))
        for remote, seed in zip(self.remotes, seeds):
            remote.send(("seed", seed))
        return [self._recv(i) for i in range(self.num_envs)]

    def close(self):
        self._assert_not_closed()
        self.closed = True
        for remote, p in zip(self.remotes, self.ps):
            if p.is_alive():
                try:
                    remote.send(("close", None))
                except (BrokenPipeError, OSError):
                    pass
        for p in self.ps:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()
        if self._shm is not None:
            if self._pinned:
                torch.cuda.cudart().cudaHostUnregister(self._slab_tensor.data_ptr())
            self._slab_tensor = None
            self._slab = None
            try:
                self._shm.close()
            except BufferError:   # an observation list handed out earlier still views the slab
                pass
            self._shm.unlink()
            self._shm = None

    @property
    def num_envs(self):
        return len(self.remotes)

    def _assert_not_closed(self):
        assert not self.closed, "This env is already closed"
