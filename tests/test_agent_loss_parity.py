"""'fp32 losses within 1e-5': the agents' _compute_loss on the reference's
fixed batch + fixed weights must reproduce the reference's loss, per-sample
priority errors and gradient norm (tests/golden/agent_losses.npz, produced by
the real reference on CPU).  Runs on CPU (torch formulation) and, with the gpu
marker, on CUDA through the fused kernels."""
import os

import numpy as np
import pytest
import torch

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "agent_losses.npz"))
OBS, NA = 10, 4


def _agent(name, device):
    from pfrl_b200 import agents, explorers, q_functions
    from pfrl_b200.replay_buffers import HostReplayBuffer

    spec = {
        "dqn": (agents.DQN, lambda: q_functions.FCStateQFunctionWithDiscreteAction(OBS, NA, 32, 2), {}),
        "ddqn": (agents.DoubleDQN,
                 lambda: q_functions.FCStateQFunctionWithDiscreteAction(OBS, NA, 32, 2),
                 dict(clip_delta=False, batch_accumulator="sum")),
        "c51": (agents.CategoricalDQN,
                lambda: q_functions.DistributionalFCStateQFunctionWithDiscreteAction(
                    OBS, NA, 51, -10, 10, 32, 2), {}),
        "rainbow": (agents.CategoricalDoubleDQN,
                    lambda: q_functions.DistributionalFCStateQFunctionWithDiscreteAction(
                        OBS, NA, 21, -2, 2, 32, 2), {}),
    }[name]
    cls, make_q, kw = spec
    qf = make_q()
    agent = cls(qf, torch.optim.SGD(qf.parameters(), lr=0.0), HostReplayBuffer(100), 0.99,
                explorers.Greedy(), gpu=0 if device == "cuda" else None, replay_start_size=10,
                minibatch_size=8, **kw)
    for which, mod in (("model", agent.model), ("target", agent.target_model)):
        sd = {k: torch.tensor(G["%s_%s_%s" % (name, which, k)]) for k in mod.state_dict()}
        mod.load_state_dict(sd)
    return agent


def _check(name, device):
    agent = _agent(name, device)
    batch = {k[len("batch_"):]: torch.tensor(G[k]).to(device) for k in G.files
             if k.startswith("batch_")}
    for use_w in (1, 0):
        eb = dict(batch)
        if not use_w:
            del eb["weights"]
        agent.model.zero_grad()
        loss, delta = agent._compute_loss(eb, want_errors=True)
        loss.backward()
        gn = torch.sqrt(sum((p.grad ** 2).sum() for p in agent.model.parameters())).item()
        np.testing.assert_allclose(loss.item(), G["%s_w%d_loss" % (name, use_w)], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(delta.cpu().numpy(), G["%s_w%d_errors" % (name, use_w)],
                                   rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(gn, G["%s_w%d_gradnorm" % (name, use_w)], rtol=1e-4)


@pytest.mark.parametrize("name", ["dqn", "ddqn", "c51", "rainbow"])
def test_cpu_torch_formulation_matches_reference(name):
    _check(name, "cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dqn", "ddqn", "c51", "rainbow"])
def test_cuda_fused_kernels_match_reference(name):
    torch.backends.cuda.matmul.allow_tf32 = False
    _check(name, "cuda")
