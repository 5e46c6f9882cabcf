"""Proximal Policy Optimization on a device-resident rollout.

Public surface = pfrl/agents/ppo.py:260-817 (constructor arguments,
act/observe, statistics names, saved_attributes).  The reference keeps the
rollout as Python lists of transition dicts, writes V(s), V(s'), log pi back
into them one float at a time and runs GAE as a Python loop per episode
(:36-53, :110-142); every minibatch rebuilds tensors from dict lists
(:488-511) and pulls two losses to the host (:662-663).

Here the rollout is a time-major set of device tensors [T, E, ...]; the
dataset is built by two batched forwards and ONE fused GAE kernel (segmented
reverse scan + advantage moments), minibatches are index gathers, and the
clipped-surrogate / value / entropy loss is one fused forward+backward kernel.
Episode segmentation (cut at done OR reset OR the flush of unfinished episodes
at update time, ppo.py:450-458,786-789) and the dataset order (episode
completion order, which fixes what ``random.sample`` draws) follow the
reference, so a seeded run visits the same minibatches.

Recurrent models are out of scope.
"""
import itertools
import random

import numpy as np
import torch
import torch.nn.functional as F

from pfrl_b200 import agent
from pfrl_b200.agents.dqn import _DeviceRing
from pfrl_b200.agents.soft_actor_critic import mode_of_distribution
from pfrl_b200.ops import ppo as fused
from pfrl_b200.utils.batch_states import batch_states
from pfrl_b200.utils.modes import evaluating, no_distribution_validation


def _elementwise_clip(x, x_min, x_max):
    return torch.min(torch.max(x, x_min), x_max)


def _yield_minibatch_indices(order, minibatch_size, num_epochs):
    """Index version of ppo.py:247-257: reshuffle the whole dataset with
    ``random.sample`` whenever the pool runs short, serve from its tail."""
    assert order
    n_total = len(order)
    pool = []
    served = 0
    while served < n_total * num_epochs:
        while len(pool) < minibatch_size:
            pool = random.sample(order, k=n_total) + pool
        yield pool[-minibatch_size:]
        served += minibatch_size
        pool = pool[:-minibatch_size]


_yield_minibatches = _yield_minibatch_indices  # the reference's name (works on any list)


class PPO(agent.AttributeSavingMixin, agent.BatchAgent):
    saved_attributes = ("model", "optimizer", "obs_normalizer")

    def __init__(self, model, optimizer, obs_normalizer=None, gpu=None, gamma=0.99, lambd=0.95,
                 phi=lambda x: x, value_func_coef=1.0, entropy_coef=0.01, update_interval=2048,
                 minibatch_size=64, epochs=10, clip_eps=0.2, clip_eps_vf=None,
                 standardize_advantages=True, batch_states=batch_states, recurrent=False,
                 max_recurrent_sequence_len=None, act_deterministically=False,
                 max_grad_norm=None, value_stats_window=1000, entropy_stats_window=1000,
                 value_loss_stats_window=100, policy_loss_stats_window=100, grad_sync=None,
                 stats_sync=None, cuda_graph=False):
        if recurrent:
            raise NotImplementedError("recurrent PPO is out of scope of pfrl_b200")
        self.model = model
        self.optimizer = optimizer
        self.obs_normalizer = obs_normalizer
        if gpu is not None and gpu >= 0:
            assert torch.cuda.is_available()
            self.device = torch.device("cuda:{}".format(gpu))
            self.model.to(self.device)
            if self.obs_normalizer is not None:
                self.obs_normalizer.to(self.device)
        else:
            self.device = torch.device("cpu")
        self.gamma = gamma
        self.lambd = lambd
        self.phi = phi
        self.value_func_coef = value_func_coef
        self.entropy_coef = entropy_coef
        self.update_interval = update_interval
        self.minibatch_size = minibatch_size
        self.epochs = epochs
        self.clip_eps = clip_eps
        self.clip_eps_vf = clip_eps_vf
        self.standardize_advantages = standardize_advantages
        self.batch_states = batch_states
        self.recurrent = False
        self.act_deterministically = act_deterministically
        self.max_grad_norm = max_grad_norm
        # data-parallel hooks: grad_sync(model) all-reduces gradients,
        # stats_sync(tensor[3] = count, sum, sumsq) all-reduces advantage moments
        self.grad_sync = grad_sync
        self.stats_sync = stats_sync
        self.use_fused = True
        # Optional: the minibatch step (gather, forward, fused loss, backward, clip,
        # Adam: ~100 launches of tiny kernels, 320 times per update) replayed as ONE
        # CUDA graph over persistent dataset buffers.
        self._graph_enabled = bool(cuda_graph) and self.device.type == "cuda" \
            and grad_sync is None
        self._graph = None
        self._graph_key = None
        if self._graph_enabled:
            for group in optimizer.param_groups:
                if "capturable" in group:
                    group["capturable"] = True
        # rollout (device tensors per vector step)
        self._states, self._next_states, self._actions = [], [], []
        self._rewards, self._nonterminal, self._cut = [], [], []
        self._segments = []     # finished episodes: (env, t_start, t_end) in completion order
        self._seg_start = None  # per env: first step of the running episode
        self._cur_state = None
        self._cur_action = None
        self.value_record = _DeviceRing(value_stats_window)
        self.entropy_record = _DeviceRing(entropy_stats_window)
        self.value_loss_record = _DeviceRing(value_loss_stats_window)
        self.policy_loss_record = _DeviceRing(policy_loss_stats_window)
        self.explained_variance = np.nan
        self.n_updates = 0

    # ----------------------------------------------------------------- acting
    def _features(self, batch_obs):
        b_state = self.batch_states(batch_obs, self.device, self.phi)
        if self.obs_normalizer:
            return b_state, self.obs_normalizer(b_state, update=False)
        return b_state, b_state

    def batch_act(self, batch_obs):
        raw, feat = self._features(batch_obs)
        with torch.no_grad(), evaluating(self.model):
            action_distrib, batch_value = self.model(feat)
            if not self.training:
                if self.act_deterministically:
                    return mode_of_distribution(action_distrib).cpu().numpy()
                return action_distrib.sample().cpu().numpy()
            action = action_distrib.sample()
            self.entropy_record.extend(action_distrib.entropy())
            self.value_record.extend(batch_value)
        self._cur_state = raw
        self._cur_action = action
        if self._seg_start is None:
            self._seg_start = [0] * len(batch_obs)
        return action.cpu().numpy()

    def batch_observe(self, batch_obs, batch_reward, batch_done, batch_reset):
        if not self.training:
            return
        assert self._cur_state is not None
        num_envs = len(batch_obs)
        next_raw = self.batch_states(batch_obs, self.device, self.phi)
        done = np.asarray(batch_done, dtype=bool)
        end = np.logical_or(done, np.asarray(batch_reset, dtype=bool))
        t = len(self._rewards)
        self._states.append(self._cur_state)
        self._actions.append(self._cur_action)
        self._next_states.append(next_raw)
        self._rewards.append(np.asarray(batch_reward, dtype=np.float32))
        self._nonterminal.append((~done).astype(np.float32))
        self._cut.append(end.astype(np.uint8))
        for i in np.nonzero(end)[0]:  # episode completion order = env order within a step
            self._segments.append((int(i), self._seg_start[i], t))
            self._seg_start[i] = t + 1
        self._cur_state = None
        self._cur_action = None
        if (t + 1) * num_envs >= self.update_interval:
            self._update_from_rollout()

    # ------------------------------------------------------------ the update
    def _dataset_order(self, T, E):
        """Flat indices t*E + e in the order of the reference's dataset list:
        finished episodes in completion order, then (flush, ppo.py:450-458)
        the unfinished ones by env index."""
        segs = list(self._segments)
        for e in range(E):
            if self._seg_start[e] <= T - 1:
                segs.append((e, self._seg_start[e], T - 1))
        order = list(itertools.chain.from_iterable(
            (range(s * E + e, (t_end + 1) * E + e, E)) for e, s, t_end in segs))
        assert len(order) == T * E
        return order

    def _update_from_rollout(self):
        T, dev = len(self._rewards), self.device
        states = torch.stack(self._states)            # [T, E, ...]
        next_states = torch.stack(self._next_states)
        actions = torch.stack(self._actions)
        E = states.shape[1]
        N = T * E
        reward = torch.as_tensor(np.stack(self._rewards), device=dev)
        nonterminal = torch.as_tensor(np.stack(self._nonterminal), device=dev)
        cut_np = np.stack(self._cut)
        cut_np[T - 1, :] = 1  # flush of unfinished episodes
        cut = torch.as_tensor(cut_np, device=dev)
        order = self._dataset_order(T, E)

        flat_s = states.reshape((N,) + states.shape[2:])
        flat_ns = next_states.reshape((N,) + states.shape[2:])
        flat_a = actions.reshape((N,) + actions.shape[2:])
        norm = self.obs_normalizer
        # _add_log_prob_and_value_to_episodes (ppo.py:110-142)
        with torch.no_grad(), evaluating(self.model):
            distribs, vs_pred = self.model(norm(flat_s, update=False) if norm else flat_s)
            _, next_vs_pred = self.model(norm(flat_ns, update=False) if norm else flat_ns)
            log_probs_old = distribs.log_prob(flat_a).float()
            v = vs_pred.reshape(T, E).float()
            v_next = next_vs_pred.reshape(T, E).float()
        # GAE (ppo.py:36-53) + advantage moments (:476-478)
        if self.use_fused and dev.type == "cuda":
            adv, v_teacher, stats = fused.gae(reward, nonterminal, v, v_next, cut, self.gamma,
                                              self.lambd)
        else:
            adv, v_teacher = _gae_torch(reward, nonterminal, v, v_next, cut, self.gamma, self.lambd)
            std, mean = torch.std_mean(adv, unbiased=False)
            stats = torch.stack([mean, std])
        if self.stats_sync is not None and self.standardize_advantages:
            stats = self.stats_sync(adv)
        adv, v_teacher = adv.reshape(N), v_teacher.reshape(N)
        v_old = v.reshape(N)

        if norm:  # _update_obs_normalizer (ppo.py:460-463), after the dataset is built
            norm.experience(flat_s)

        adv_stats = stats if self.standardize_advantages else None
        if self._graph_enabled:
            self._minibatches_graphed(order, flat_s, flat_a, log_probs_old, v_old, adv, v_teacher,
                                      adv_stats)
        else:
            for mb in _yield_minibatch_indices(order, self.minibatch_size, self.epochs):
                idx = torch.as_tensor(mb, device=dev)
                s = flat_s[idx]
                self._minibatch_step(norm(s, update=False) if norm else s, flat_a[idx],
                                     log_probs_old[idx], v_old[idx], adv[idx], v_teacher[idx],
                                     adv_stats)
                self.n_updates += 1

        # explained variance of the value predictions (ppo.py:56-62)
        var_t = torch.var(v_teacher, unbiased=False)
        ev = 1 - torch.var(v_teacher - v_old, unbiased=False) / var_t
        self._explained_variance_t = torch.where(var_t == 0, torch.full_like(ev, float("nan")), ev)
        self.explained_variance = None
        self._states, self._next_states, self._actions = [], [], []
        self._rewards, self._nonterminal, self._cut = [], [], []
        self._segments = []
        self._seg_start = [0] * E

    def _minibatch_step(self, s, a, log_probs_old, v_old, advs, v_teacher, adv_stats,
                        record=True):
        """One gradient step on one minibatch (ppo.py:480-532)."""
        distribs, vs_pred = self.model(s)
        self.model.zero_grad()
        loss = self._lossfun(
            distribs.entropy(), vs_pred, distribs.log_prob(a), vs_pred_old=v_old[..., None],
            log_probs_old=log_probs_old, advs=advs, vs_teacher=v_teacher[..., None],
            adv_stats=adv_stats, record=record)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync(self.model)
        if self.max_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), self.max_grad_norm)
        self.optimizer.step()

    def _minibatches_graphed(self, order, flat_s, flat_a, log_probs_old, v_old, adv, v_teacher,
                             adv_stats):
        dev, norm = self.device, self.obs_normalizer
        # the normaliser is frozen during the epochs (update=False): apply it once
        data = dict(s=norm(flat_s, update=False) if norm else flat_s, a=flat_a,
                    lp=log_probs_old, vo=v_old, adv=adv, vt=v_teacher)
        if adv_stats is not None:
            data["stats"] = adv_stats
        mbs = list(_yield_minibatch_indices(order, self.minibatch_size, self.epochs))
        all_idx = torch.as_tensor(np.asarray(mbs, dtype=np.int64), device=dev)
        key = tuple((k, tuple(v.shape), v.dtype) for k, v in data.items()) + (self.minibatch_size,)
        if self._graph_key != key:  # (re)allocate the persistent buffers, drop the old graph
            self._graph, self._graph_key = None, key
            self._g = {k: torch.empty_like(v) for k, v in data.items()}
            self._g_idx = torch.zeros(self.minibatch_size, dtype=torch.int64, device=dev)
            self._g_warm = 0
        for k, v in data.items():
            self._g[k].copy_(v)
        g = self._g

        def step(record):
            i = self._g_idx
            self._minibatch_step(g["s"][i], g["a"][i], g["lp"][i], g["vo"][i], g["adv"][i],
                                 g["vt"][i], g.get("stats"), record=record)

        for row in range(all_idx.shape[0]):
            self._g_idx.copy_(all_idx[row])
            if self._graph is None and self._g_warm < 3:
                self._g_warm += 1
                step(record=True)
            else:
                if self._graph is None:
                    torch.cuda.synchronize(dev)
                    self._graph = torch.cuda.CUDAGraph()
                    with no_distribution_validation(), torch.cuda.graph(self._graph):
                        step(record=False)
                # capture only records the work: every row (this one included) is replayed
                self._graph.replay()
                self.policy_loss_record.append(self._last_loss_parts[1])
                self.value_loss_record.append(self._last_loss_parts[2])
            self.n_updates += 1

    def _lossfun(self, entropy, vs_pred, log_probs, vs_pred_old, log_probs_old, advs, vs_teacher,
                 adv_stats=None, record=True):
        """ppo.py:495 + :634-671."""
        if self.use_fused and log_probs.is_cuda:
            loss, parts = fused.ppo_loss(
                log_probs, entropy, vs_pred, log_probs_old, vs_pred_old, advs, vs_teacher,
                adv_stats, self.clip_eps, self.clip_eps_vf, self.value_func_coef,
                self.entropy_coef)
            self._last_loss_parts = parts  # static tensor when graph-captured
            if record:
                self.policy_loss_record.append(parts[1])
                self.value_loss_record.append(parts[2])
            return loss
        if adv_stats is not None:
            advs = (advs - adv_stats[0]) / (adv_stats[1] + 1e-8)
        prob_ratio = torch.exp(log_probs - log_probs_old)
        loss_policy = -torch.mean(torch.min(
            prob_ratio * advs,
            torch.clamp(prob_ratio, 1 - self.clip_eps, 1 + self.clip_eps) * advs))
        if self.clip_eps_vf is None:
            loss_value_func = F.mse_loss(vs_pred, vs_teacher)
        else:
            clipped = _elementwise_clip(vs_pred, vs_pred_old - self.clip_eps_vf,
                                        vs_pred_old + self.clip_eps_vf)
            loss_value_func = torch.mean(torch.max(
                F.mse_loss(vs_pred, vs_teacher, reduction="none"),
                F.mse_loss(clipped, vs_teacher, reduction="none")))
        loss_entropy = -torch.mean(entropy)
        self._last_loss_parts = torch.stack(
            [loss_policy.detach() * 0, loss_policy.detach(), loss_value_func.detach()])
        if record:
            self.value_loss_record.append(loss_value_func.detach())
            self.policy_loss_record.append(loss_policy.detach())
        return (loss_policy + self.value_func_coef * loss_value_func
                + self.entropy_coef * loss_entropy)

    def get_statistics(self):
        if self.explained_variance is None:
            self.explained_variance = float(self._explained_variance_t.item())
        return [
            ("average_value", self.value_record.mean()),
            ("average_entropy", self.entropy_record.mean()),
            ("average_value_loss", self.value_loss_record.mean()),
            ("average_policy_loss", self.policy_loss_record.mean()),
            ("n_updates", self.n_updates),
            ("explained_variance", self.explained_variance),
        ]


def _gae_torch(reward, nonterminal, v, v_next, cut, gamma, lambd):
    """Host formulation of the GAE recurrence (CPU agents; fp64 like the kernel)."""
    T, E = reward.shape
    adv = torch.zeros((T, E), dtype=torch.float64, device=reward.device)
    run = torch.zeros(E, dtype=torch.float64, device=reward.device)
    r, nt = reward.double(), nonterminal.double()
    vv, vn = v.double(), v_next.double()
    for t in range(T - 1, -1, -1):
        run = torch.where(cut[t].bool(), torch.zeros_like(run), run)
        run = (r[t] + gamma * nt[t] * vn[t] - vv[t]) + gamma * lambd * run
        adv[t] = run
    return adv.float(), (adv + vv).float()
