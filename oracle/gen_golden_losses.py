"""Golden vectors for the loss / GAE kernels, produced by the REAL reference
(called from oracle/gen_golden.py; TEST INFRASTRUCTURE ONLY)."""
import os

import numpy as np


def main(out_dir):
    import torch
    from pfrl.agents import categorical_dqn, dqn, iqn, ppo

    rng = np.random.RandomState(2024)
    g = {}

    # ---- scalar TD loss (dqn.py:44-104, 388-470) --------------------------------
    B, nA = 96, 18
    q = rng.randn(B, nA).astype(np.float32) * 2
    act = rng.randint(0, nA, size=B)
    next_q = rng.randn(B).astype(np.float32) * 2
    rew = rng.choice([-1.0, 0.0, 1.0, 2.5], size=B).astype(np.float32)
    disc = (0.99 ** rng.randint(1, 4, size=B)).astype(np.float32)
    term = (rng.rand(B) < 0.2).astype(np.float32)
    w = (rng.rand(B) + 0.1).astype(np.float32)
    g.update(td_q=q, td_action=act, td_next_q=next_q, td_reward=rew, td_discount=disc,
             td_terminal=term, td_weights=w)
    for clip in (True, False):
        for acc in ("mean", "sum"):
            for use_w in (True, False):
                qt = torch.tensor(q, requires_grad=True)
                y = qt[torch.arange(B), torch.tensor(act)]
                t = torch.tensor(rew) + torch.tensor(disc) * (1.0 - torch.tensor(term)) * \
                    torch.tensor(next_q)
                if use_w:
                    loss = dqn.compute_weighted_value_loss(y, t, torch.tensor(w), clip, acc)
                else:
                    loss = dqn.compute_value_loss(y, t, clip, acc)
                loss.backward()
                key = "td_%d_%s_%d" % (clip, acc, use_w)
                g[key + "_loss"] = loss.detach().numpy()
                g[key + "_grad"] = qt.grad.numpy()
    g["td_delta"] = np.abs(q[np.arange(B), act] - (rew + disc * (1 - term) * next_q))

    # ---- C51 (categorical_dqn.py:7-57, 60-97, 178-204) --------------------------
    B, n = 80, 51
    z = torch.linspace(-10, 10, n, dtype=torch.float32)
    logits = rng.randn(B, n).astype(np.float32)
    y = torch.softmax(torch.tensor(logits), dim=1)
    next_p = torch.softmax(torch.tensor(rng.randn(B, n).astype(np.float32) * 2), dim=1)
    rew = rng.choice([-1.0, 0.0, 1.0, 3.0, -12.0, 15.0], size=B).astype(np.float32)
    disc = (0.99 ** rng.randint(1, 4, size=B)).astype(np.float32)
    term = (rng.rand(B) < 0.25).astype(np.float32)
    w = (rng.rand(B) + 0.1).astype(np.float32)
    Tz = (torch.tensor(rew)[..., None] + (1.0 - torch.tensor(term)[..., None])
          * torch.tensor(disc)[..., None] * z[None])
    t = categorical_dqn._apply_categorical_projection(Tz, next_p, z)
    g.update(c51_z=z.numpy(), c51_y=y.numpy(), c51_next_p=next_p.numpy(), c51_reward=rew,
             c51_discount=disc, c51_terminal=term, c51_weights=w, c51_target=t.numpy())
    for acc in ("mean", "sum"):
        for use_w in (True, False):
            yt = y.clone().requires_grad_(True)
            elt = -t * torch.log(torch.clamp(yt, 1e-10, 1.0))
            if use_w:
                loss = categorical_dqn.compute_weighted_value_loss(elt, B, torch.tensor(w), acc)
            else:
                loss = categorical_dqn.compute_value_loss(elt, acc)
            loss.backward()
            key = "c51_%s_%d" % (acc, use_w)
            g[key + "_loss"] = loss.detach().numpy()
            g[key + "_grad"] = yt.grad.numpy()
            g["c51_delta"] = elt.detach().sum(dim=1).numpy()
    # a hand-checkable projection case incl. inexact delta_z and out-of-range atoms
    z2 = torch.linspace(-1, 1, 7, dtype=torch.float32)
    yv = torch.tensor([[-3.0, -1.0, -0.2, 0.0, 1.0 / 3, 0.999, 5.0]], dtype=torch.float32)
    pv = torch.tensor([[0.1, 0.2, 0.05, 0.25, 0.1, 0.2, 0.1]], dtype=torch.float32)
    g.update(proj_z=z2.numpy(), proj_y=yv.numpy(), proj_p=pv.numpy(),
             proj_out=categorical_dqn._apply_categorical_projection(yv, pv, z2).numpy())

    # ---- quantile Huber (iqn.py:176-250) ----------------------------------------
    B, N, Np = 24, 16, 12
    yq = rng.randn(B, N).astype(np.float32)
    tq = (rng.randn(B, Np) * 1.5).astype(np.float32)
    taus = rng.rand(B, N).astype(np.float32)
    w = (rng.rand(B) + 0.1).astype(np.float32)
    g.update(qh_y=yq, qh_t=tq, qh_taus=taus, qh_weights=w)
    for acc in ("mean", "sum"):
        for use_w in (True, False):
            yt = torch.tensor(yq, requires_grad=True)
            elt = iqn.compute_eltwise_huber_quantile_loss(yt, torch.tensor(tq), torch.tensor(taus))
            if use_w:
                loss = iqn.compute_weighted_value_loss(elt, torch.tensor(w), acc)
            else:
                loss = iqn.compute_value_loss(elt, acc)
            loss.backward()
            key = "qh_%s_%d" % (acc, use_w)
            g[key + "_loss"] = loss.detach().numpy()
            g[key + "_grad"] = yt.grad.numpy()
            g["qh_delta"] = elt.detach().mean((1, 2)).numpy()

    # ---- GAE (ppo.py:36-53): episodes of Python-float transitions --------------
    T, E = 40, 6
    rew = rng.randn(T, E)
    v = rng.randn(T, E).astype(np.float32)
    v_next = rng.randn(T, E).astype(np.float32)
    nonterm = (rng.rand(T, E) > 0.1).astype(np.float64)
    cut = (nonterm == 0) | (rng.rand(T, E) < 0.05)
    cut[-1, :] = True
    adv = np.zeros((T, E))
    vt = np.zeros((T, E))
    for gamma, lambd, tag in ((0.995, 0.95, "a"), (0.9, 0.5, "b"), (1.0, 1.0, "c")):
        for e in range(E):
            start = 0
            for tt in range(T):
                if cut[tt, e]:
                    ep = [dict(reward=float(rew[k, e]), nonterminal=float(nonterm[k, e]),
                               v_pred=float(v[k, e]), next_v_pred=float(v_next[k, e]))
                          for k in range(start, tt + 1)]
                    ppo._add_advantage_and_value_target_to_episode(ep, gamma, lambd)
                    for k, tr in zip(range(start, tt + 1), ep):
                        adv[k, e] = tr["adv"]
                        vt[k, e] = tr["v_teacher"]
                    start = tt + 1
        g["gae_%s_adv" % tag] = adv.copy()
        g["gae_%s_vt" % tag] = vt.copy()
        g["gae_%s_params" % tag] = np.array([gamma, lambd])
    g.update(gae_reward=rew, gae_v=v, gae_v_next=v_next, gae_nonterminal=nonterm, gae_cut=cut)

    # ---- PPO loss (ppo.py:495, 634-671) through a real PPO instance -------------
    M = 200
    lp = (rng.randn(M) * 0.3 - 1.0).astype(np.float32)
    lp_old = (lp + rng.randn(M).astype(np.float32) * 0.25).astype(np.float32)
    ent = (rng.rand(M) + 0.5).astype(np.float32)
    vp = rng.randn(M, 1).astype(np.float32)
    vp_old = (vp + rng.randn(M, 1).astype(np.float32) * 0.3).astype(np.float32)
    vteach = rng.randn(M, 1).astype(np.float32)
    advs = rng.randn(M).astype(np.float32) * 2 + 0.3
    g.update(ppo_lp=lp, ppo_lp_old=lp_old, ppo_ent=ent, ppo_v=vp, ppo_v_old=vp_old,
             ppo_vt=vteach, ppo_adv=advs)
    model = torch.nn.Linear(2, 2)
    for clip_vf, tag in ((None, "a"), (0.2, "b")):
        agent = ppo.PPO(model, torch.optim.SGD(model.parameters(), lr=0.1), clip_eps=0.2,
                        clip_eps_vf=clip_vf, value_func_coef=0.5, entropy_coef=0.01)
        all_advs = torch.tensor(advs)
        std_a, mean_a = torch.std_mean(all_advs, unbiased=False)
        a_n = (all_advs - mean_a) / (std_a + 1e-8)
        t_lp = torch.tensor(lp, requires_grad=True)
        t_ent = torch.tensor(ent, requires_grad=True)
        t_v = torch.tensor(vp, requires_grad=True)
        loss = agent._lossfun(t_ent, t_v, t_lp, vs_pred_old=torch.tensor(vp_old),
                              log_probs_old=torch.tensor(lp_old), advs=a_n,
                              vs_teacher=torch.tensor(vteach))
        loss.backward()
        g["ppo_%s_loss" % tag] = loss.detach().numpy()
        g["ppo_%s_policy" % tag] = np.float32(agent.policy_loss_record[-1])
        g["ppo_%s_value" % tag] = np.float32(agent.value_loss_record[-1])
        g["ppo_%s_g_lp" % tag] = t_lp.grad.numpy()
        g["ppo_%s_g_ent" % tag] = t_ent.grad.numpy()
        g["ppo_%s_g_v" % tag] = t_v.grad.numpy()
        g["ppo_mean_std"] = np.array([float(mean_a), float(std_a)], dtype=np.float32)

    agent_losses(out_dir)
    normalizer(out_dir)
    a2c_trace(out_dir)
    agent_traces(out_dir)
    more_agent_traces(out_dir)
    driver_trace(out_dir)
    g["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(out_dir, "losses.npz"), **g)
    print("wrote losses.npz with", len(g), "arrays")


def agent_losses(out_dir):
    """End-to-end _compute_loss of the reference's agents on a fixed batch and
    fixed weights (fp32, CPU): pins 'fp32 losses within 1e-5'."""
    import torch

    from pfrl import agents, explorers, q_functions, replay_buffers

    rng = np.random.RandomState(7)
    torch.manual_seed(7)
    B, obs, nA = 48, 10, 4
    batch = dict(
        state=rng.randn(B, obs).astype(np.float32), next_state=rng.randn(B, obs).astype(np.float32),
        action=rng.randint(0, nA, size=B).astype(np.int64),
        reward=rng.choice([-1.0, 0.0, 1.0, 0.3], size=B).astype(np.float32),
        discount=(0.99 ** rng.randint(1, 4, size=B)).astype(np.float32),
        is_state_terminal=(rng.rand(B) < 0.2).astype(np.float32),
        weights=(rng.rand(B) + 0.2).astype(np.float32))
    g = {"batch_" + k: v for k, v in batch.items()}
    tb = {k: torch.tensor(v) for k, v in batch.items()}

    def run(name, cls, qf, **kw):
        target_init = None
        agent = cls(qf, torch.optim.SGD(qf.parameters(), lr=0.0),
                    replay_buffers.PrioritizedReplayBuffer(100), 0.99, explorers.Greedy(),
                    replay_start_size=10, minibatch_size=8, **kw)
        # make the target net differ from the online net
        with torch.no_grad():
            for p in agent.target_model.parameters():
                p.add_(torch.randn_like(p) * 0.05)
        for k, v in agent.model.state_dict().items():
            g["%s_model_%s" % (name, k)] = v.numpy().copy()
        for k, v in agent.target_model.state_dict().items():
            g["%s_target_%s" % (name, k)] = v.numpy().copy()
        for use_w in (1, 0):
            eb = dict(tb)
            if not use_w:
                del eb["weights"]
            errs = []
            agent.model.zero_grad()
            loss = agent._compute_loss(eb, errors_out=errs)
            loss.backward()
            g["%s_w%d_loss" % (name, use_w)] = loss.detach().numpy()
            g["%s_w%d_errors" % (name, use_w)] = np.asarray(errs, dtype=np.float32)
            g["%s_w%d_gradnorm" % (name, use_w)] = np.float32(
                torch.sqrt(sum((p.grad ** 2).sum() for p in agent.model.parameters())))

    run("dqn", agents.DQN, q_functions.FCStateQFunctionWithDiscreteAction(obs, nA, 32, 2))
    run("ddqn", agents.DoubleDQN, q_functions.FCStateQFunctionWithDiscreteAction(obs, nA, 32, 2),
        clip_delta=False, batch_accumulator="sum")
    run("c51", agents.CategoricalDQN,
        q_functions.DistributionalFCStateQFunctionWithDiscreteAction(obs, nA, 51, -10, 10, 32, 2))
    run("rainbow", agents.CategoricalDoubleDQN,
        q_functions.DistributionalFCStateQFunctionWithDiscreteAction(obs, nA, 21, -2, 2, 32, 2))
    np.savez_compressed(os.path.join(out_dir, "agent_losses.npz"), **g)
    print("wrote agent_losses.npz with", len(g), "arrays")


def normalizer(out_dir):
    """EmpiricalNormalization (pfrl/nn/empirical_normalization.py:6-109) on a
    fixed sequence of batches."""
    import torch
    from pfrl.nn import EmpiricalNormalization

    rng = np.random.RandomState(3)
    en = EmpiricalNormalization(7, clip_threshold=5)
    xs = [(rng.randn(n, 7) * s + m).astype(np.float32)
          for n, s, m in ((16, 1.0, 0.0), (1, 3.0, 2.0), (64, 0.5, -1.0), (5, 10.0, 4.0))]
    outs = [en(torch.tensor(x), update=True).numpy() for x in xs]
    probe = rng.randn(9, 7).astype(np.float32) * 4
    np.savez_compressed(
        os.path.join(out_dir, "empirical_normalization.npz"),
        **{"x%d" % i: x for i, x in enumerate(xs)}, **{"y%d" % i: y for i, y in enumerate(outs)},
        probe=probe, probe_out=en(torch.tensor(probe), update=False).numpy(),
        mean=en.mean.numpy(), std=en.std.numpy(), count=np.int64(en.count.item()))
    print("wrote empirical_normalization.npz")


def a2c_trace(out_dir):
    """A seeded A2C run of the reference (pfrl/agents/a2c.py:14-310) on scripted
    observations / rewards / dones: the sampled actions, the three running
    statistics after every update and the final parameters."""
    import torch
    from torch import nn

    import pfrl
    from pfrl.agents import a2c
    from pfrl.policies import GaussianHeadWithStateIndependentCovariance, SoftmaxCategoricalHead

    rng = np.random.RandomState(5)
    N, obs, steps = 6, 9, 23
    g = dict(obs=rng.randn(steps + 1, N, obs).astype(np.float32),
             reward=rng.randn(steps, N).astype(np.float32),
             done=(rng.rand(steps, N) < 0.15))

    def make(kind):
        torch.manual_seed(31)
        if kind == "discrete":
            head = nn.Sequential(nn.Linear(16, 4), SoftmaxCategoricalHead())
        else:
            head = nn.Sequential(nn.Linear(16, 3), GaussianHeadWithStateIndependentCovariance(
                action_size=3, var_type="diagonal", var_func=lambda x: torch.exp(2 * x),
                var_param_init=0))
        return nn.Sequential(nn.Linear(obs, 16), nn.Tanh(),
                             pfrl.nn.Branched(head, nn.Linear(16, 1)))

    for kind, kw in (("discrete", dict(use_gae=False, max_grad_norm=0.5)),
                     ("gaussian", dict(use_gae=True, tau=0.9, max_grad_norm=None))):
        model = make(kind)
        for k, v in model.state_dict().items():
            g["%s_init_%s" % (kind, k)] = v.numpy().copy()
        opt = torch.optim.RMSprop(model.parameters(), lr=7e-3, eps=1e-5, alpha=0.99)
        agent = a2c.A2C(model, opt, gamma=0.97, num_processes=N, update_steps=4,
                        average_actor_loss_decay=0.0, average_entropy_decay=0.0,
                        average_value_decay=0.0, **kw)
        torch.manual_seed(77)
        actions, stats = [], []
        for t in range(steps):
            actions.append(agent.batch_act(list(g["obs"][t])))
            agent.batch_observe(list(g["obs"][t + 1]), list(g["reward"][t]), list(g["done"][t]),
                                [False] * N)
            stats.append([v for _, v in agent.get_statistics()])
        g[kind + "_actions"] = np.asarray(actions)
        g[kind + "_stats"] = np.asarray(stats, dtype=np.float64)
        for k, v in model.state_dict().items():
            g["%s_final_%s" % (kind, k)] = v.numpy().copy()
    np.savez_compressed(os.path.join(out_dir, "a2c_trace.npz"), **g)
    print("wrote a2c_trace.npz with", len(g), "arrays")


def _make_trace_agent(lib, kind, rbuf):
    """The same construction code serves the reference (lib = pfrl) and the
    rebuild (lib = pfrl_b200) -- tests/test_agent_traces_cpu.py imports it."""
    import torch

    phi = lambda x: x.astype(np.float32, copy=False)  # noqa: E731
    eps = lib.explorers.LinearDecayEpsilonGreedy(1.0, 0.1, 150, lambda: np.random.randint(2))
    common = dict(replay_start_size=40, minibatch_size=16, update_interval=1,
                  target_update_interval=20, phi=phi)
    if kind == "ddqn":
        q = lib.q_functions.FCStateQFunctionWithDiscreteAction(5, 2, 32, 2)
        cls, explorer, kw = lib.agents.DoubleDQN, eps, {}
    elif kind == "dqn_sum":
        q = lib.q_functions.FCStateQFunctionWithDiscreteAction(5, 2, 32, 2)
        cls, explorer = lib.agents.DQN, eps
        kw = dict(clip_delta=False, batch_accumulator="sum", target_update_method="soft",
                  soft_update_tau=0.05, max_grad_norm=1.0)
    elif kind == "c51":
        q = lib.q_functions.DistributionalFCStateQFunctionWithDiscreteAction(
            5, 2, 21, -1.0, 2.0, 32, 2)
        cls, explorer, kw = lib.agents.CategoricalDQN, eps, {}
    elif kind == "rainbow":
        q = lib.q_functions.DistributionalFCStateQFunctionWithDiscreteAction(
            5, 2, 21, -1.0, 2.0, 32, 2)
        lib.nn.to_factorized_noisy(q, sigma_scale=0.5)
        cls, explorer, kw = lib.agents.CategoricalDoubleDQN, lib.explorers.Greedy(), {}
    else:
        raise ValueError(kind)
    opt = torch.optim.Adam(q.parameters(), lr=1e-3, eps=1e-4)
    return q, cls(q, opt, rbuf, 0.95, explorer, **common, **kw)


def _run_trace(agent, rbuf, steps=260, n_envs=2, check=None):
    from pfrl_b200.envs import ChainEnv

    envs = [ChainEnv(seed=i) for i in range(n_envs)]
    obs = [e.reset() for e in envs]
    actions, stats = [], []
    for t in range(steps):
        a = [int(x) for x in agent.batch_act(obs)]
        if check is not None:
            check(t, a)
        actions.append(a)
        nobs, r, d, info = zip(*[e.step(x) for e, x in zip(envs, a)])
        resets = [i["needs_reset"] for i in info]
        agent.batch_observe(list(nobs), list(r), list(d), resets)
        obs = [e.reset() if (dd or rr) else o for e, o, dd, rr in zip(envs, nobs, d, resets)]
        st = dict(agent.get_statistics())
        stats.append([st["average_q"], st["average_loss"], st["n_updates"], st["rlen"]])
    return np.asarray(actions), np.asarray(stats, dtype=np.float64)


TRACE_KINDS = ("ddqn", "dqn_sum", "c51", "rainbow")
TRACE_PER = dict(alpha=0.6, beta0=0.4, betasteps=100, num_steps=3, normalize_by_max="memory")


def agent_traces(out_dir):
    """Seeded training runs of the reference's DQN family with 3-step prioritised
    replay on two chain environments: chosen actions, statistics after every
    vector step, final parameters and the final state of the priority trees.

    One adapter sits at the documented dtype boundary (DESIGN.md section 1,
    "dtype contract"): TD errors enter ``update_errors`` as Python floats.
    Unadapted, numpy >= 2 keeps them np.float32 and the reference's tree then
    sums in float32 -- a numpy-version artefact, not part of the algorithm."""
    import torch

    import pfrl

    for kind in TRACE_KINDS:
        rbuf = pfrl.replay_buffers.PrioritizedReplayBuffer(150, **TRACE_PER)
        raw_update = rbuf.update_errors
        rbuf.update_errors = lambda errors, raw=raw_update: raw([float(e) for e in errors])
        torch.manual_seed(3)
        q, agent = _make_trace_agent(pfrl, kind, rbuf)
        g = {"init_" + k: v.numpy().copy() for k, v in q.state_dict().items()}
        np.random.seed(9)
        torch.manual_seed(9)
        actions, stats = _run_trace(agent, rbuf)
        g.update(actions=actions, stats=stats,
                 total=np.float64(rbuf.memory.priority_sums.sum()),
                 min=np.float64(rbuf.memory.priority_mins.min()),
                 max_priority=np.float64(rbuf.memory.max_priority), beta=np.float64(rbuf.beta))
        for k, v in q.state_dict().items():
            g["final_" + k] = v.numpy().copy()
        np.savez_compressed(os.path.join(out_dir, "agent_trace_%s.npz" % kind), **g)
        print("wrote agent_trace_%s.npz:" % kind, int(stats[-1][2]), "updates")


# ---------------------------------------------------------------------------
# IQN, SAC, TD3, DDPG (uniform replay) and PPO
# ---------------------------------------------------------------------------
MORE_TRACE_KINDS = ("iqn", "sac", "td3", "ddpg", "ppo")


def _make_more_agent(lib, kind, rbuf, gpu=None):
    import torch
    from torch import distributions, nn

    phi = lambda x: x.astype(np.float32, copy=False)  # noqa: E731
    burn = lambda: np.random.uniform(-1, 1, size=1).astype(np.float32)  # noqa: E731

    def qf():
        return nn.Sequential(lib.nn.ConcatObsAndAction(), nn.Linear(6, 32), nn.ReLU(),
                             nn.Linear(32, 1))

    def det_policy():
        return nn.Sequential(nn.Linear(5, 32), nn.ReLU(), nn.Linear(32, 1), nn.Tanh(),
                             lib.policies.DeterministicHead())

    adam = lambda m: torch.optim.Adam(m.parameters(), lr=3e-3)  # noqa: E731
    if kind == "iqn":
        from importlib import import_module

        iqn = import_module(lib.__name__ + ".agents.iqn")
        q = iqn.ImplicitQuantileQFunction(
            psi=nn.Sequential(nn.Linear(5, 24), nn.ReLU()),
            phi=nn.Sequential(iqn.CosineBasisLinear(8, 24), nn.ReLU()),
            f=nn.Linear(24, 2))
        eps = lib.explorers.LinearDecayEpsilonGreedy(1.0, 0.1, 150, lambda: np.random.randint(2))
        return lib.agents.IQN(
            q, adam(q), rbuf, 0.95, eps, replay_start_size=40, minibatch_size=16,
            update_interval=1, target_update_interval=20, phi=phi,
            quantile_thresholds_N=6, quantile_thresholds_N_prime=5, quantile_thresholds_K=4, gpu=gpu)
    if kind == "sac":
        def squashed(x):
            mean, log_scale = torch.chunk(x, 2, dim=1)
            base = distributions.Independent(
                distributions.Normal(mean, torch.exp(torch.clamp(log_scale, -5, 2))), 1)
            return distributions.transformed_distribution.TransformedDistribution(
                base, [distributions.transforms.TanhTransform(cache_size=1)])

        policy = nn.Sequential(nn.Linear(5, 32), nn.ReLU(), nn.Linear(32, 2),
                               lib.nn.Lambda(squashed))
        q1, q2 = qf(), qf()
        return lib.agents.SoftActorCritic(
            policy, q1, q2, adam(policy), adam(q1), adam(q2), rbuf, gamma=0.95,
            replay_start_size=40, minibatch_size=16, entropy_target=-1.0,
            temperature_optimizer_lr=3e-3, phi=phi, burnin_action_func=burn, gpu=gpu)
    ex = lib.explorers.AdditiveGaussian(scale=0.3, low=-1, high=1)
    if kind == "td3":
        p, q1, q2 = det_policy(), qf(), qf()
        return lib.agents.TD3(p, q1, q2, adam(p), adam(q1), adam(q2), rbuf, 0.95, ex,
                              replay_start_size=40, minibatch_size=16, phi=phi,
                              burnin_action_func=burn, gpu=gpu)
    if kind == "ddpg":
        p, q = det_policy(), qf()
        return lib.agents.DDPG(p, q, adam(p), adam(q), rbuf, 0.95, ex, replay_start_size=40,
                               minibatch_size=16, phi=phi, target_update_method="soft",
                               target_update_interval=1, soft_update_tau=0.05,
                               burnin_action_func=burn, gpu=gpu)
    if kind == "ppo":
        model = nn.Sequential(nn.Linear(5, 32), nn.Tanh(), lib.nn.Branched(
            nn.Sequential(nn.Linear(32, 2), lib.policies.SoftmaxCategoricalHead()),
            nn.Linear(32, 1)))
        return lib.agents.PPO(
            model, torch.optim.Adam(model.parameters(), lr=3e-3),
            obs_normalizer=lib.nn.EmpiricalNormalization(5, clip_threshold=5), gamma=0.95,
            lambd=0.9, phi=phi, update_interval=64, minibatch_size=16, epochs=3, clip_eps=0.2,
            clip_eps_vf=0.3, entropy_coef=0.01, max_grad_norm=0.5, gpu=gpu)
    raise ValueError(kind)


def _module_attrs(agent):
    import torch

    return [(name, getattr(agent, name)) for name in agent.saved_attributes
            if isinstance(getattr(agent, name, None), torch.nn.Module)]


def _run_more_trace(agent, kind, steps, check=None):
    from pfrl_b200.envs import ChainEnv

    cont = kind in ("sac", "td3", "ddpg")
    envs = [ChainEnv(continuous=cont, seed=i) for i in range(2)]
    obs = [e.reset() for e in envs]
    actions, stats = [], []
    for t in range(steps):
        a = np.asarray(agent.batch_act(obs))
        if check is not None:
            check(t, a)
        actions.append(a.copy())
        nobs, r, d, info = zip(*[e.step(x) for e, x in zip(envs, a)])
        resets = [i["needs_reset"] for i in info]
        agent.batch_observe(list(nobs), list(r), list(d), resets)
        obs = [e.reset() if (dd or rr) else o for e, o, dd, rr in zip(envs, nobs, d, resets)]
        stats.append([float(v) for _, v in agent.get_statistics()])
    return np.asarray(actions), np.asarray(stats, dtype=np.float64)


def more_agent_traces(out_dir):
    """Seeded runs of the reference's IQN / SAC / TD3 / DDPG (uniform replay,
    capacity wrap-around) and PPO on two chain environments."""
    import random

    import torch

    import pfrl

    for kind in MORE_TRACE_KINDS:
        rbuf = None if kind == "ppo" else pfrl.replay_buffers.ReplayBuffer(150)
        torch.manual_seed(3)
        agent = _make_more_agent(pfrl, kind, rbuf)
        g = {}
        for name, mod in _module_attrs(agent):
            for k, v in mod.state_dict().items():
                g["init_%s__%s" % (name, k)] = v.numpy().copy()
        np.random.seed(9)
        torch.manual_seed(9)
        random.seed(9)
        actions, stats = _run_more_trace(agent, kind, 200 if kind == "ppo" else 160)
        g.update(actions=actions, stats=stats,
                 stat_names=np.array([n for n, _ in agent.get_statistics()]))
        for name, mod in _module_attrs(agent):
            for k, v in mod.state_dict().items():
                g["final_%s__%s" % (name, k)] = v.numpy().copy()
        np.savez_compressed(os.path.join(out_dir, "agent_trace_%s.npz" % kind), **g)
        print("wrote agent_trace_%s.npz" % kind, dict(zip(g["stat_names"], stats[-1])))


def _run_driver(lib, rbuf, outdir):
    """train_agent_batch_with_evaluation on two chain envs + a separate eval
    vector env: the whole loop (per-env step counting, resets, evaluator
    schedule, scores.txt) in one seeded run.  Shared by both sides."""
    import logging

    import torch

    from pfrl_b200.envs import ChainEnv

    torch.manual_seed(3)
    q, agent = _make_trace_agent(lib, "ddqn", rbuf)
    np.random.seed(9)
    torch.manual_seed(9)
    env = lib.envs.SerialVectorEnv([ChainEnv(seed=i) for i in range(2)])
    eval_env = lib.envs.SerialVectorEnv([ChainEnv(seed=10 + i) for i in range(3)])
    log = logging.getLogger("b2rl.driver_trace")
    log.setLevel(logging.CRITICAL)
    agent2, history = lib.experiments.train_agent_batch_with_evaluation(
        agent=agent, env=env, steps=700, eval_n_steps=None, eval_n_episodes=4, eval_interval=150,
        outdir=outdir, eval_env=eval_env, max_episode_len=25, log_interval=None, logger=log)
    rows = [ln.split("\t") for ln in open(os.path.join(outdir, "scores.txt")).read().strip()
            .splitlines()]
    return q, agent, history, rows


def driver_trace(out_dir):
    import tempfile

    import pfrl

    rbuf = pfrl.replay_buffers.PrioritizedReplayBuffer(150, **TRACE_PER)
    raw_update = rbuf.update_errors
    rbuf.update_errors = lambda errors: raw_update([float(e) for e in errors])
    with tempfile.TemporaryDirectory() as d:
        q, agent, history, rows = _run_driver(pfrl, rbuf, d)
        saved = sorted(os.listdir(d))
    header, body = rows[0], rows[1:]
    keep = [i for i, name in enumerate(header) if name != "elapsed"]
    g = dict(header=np.array([header[i] for i in keep]),
             scores=np.array([[float(r[i]) for i in keep] for r in body], dtype=np.float64),
             eval_scores=np.array([h["eval_score"] for h in history], dtype=np.float64),
             saved=np.array(saved), t=np.int64(agent.t), optim_t=np.int64(agent.optim_t))
    for k, v in q.state_dict().items():
        g["final_" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(out_dir, "driver_trace.npz"), **g)
    print("wrote driver_trace.npz:", len(body), "evaluations, saved", saved)
