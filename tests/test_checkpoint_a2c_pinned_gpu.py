"""GPU variants of the CPU tests (tests/fake_store.py) for the reference-pickle
restore into the CUDA store, collections.PrioritizedBuffer over the
device trees (a1-a3), the pinned observation slab of MultiprocessVectorEnv (8f2)
and A2C on a CUDA device (8f4).  First run on a B200 in round 2 (4 passed).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_pickle_into_cuda_store():
    from pfrl_b200.replay_buffer import batch_experiences
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer, ReplayBuffer
    from pfrl_b200.replay_buffers.device_buffer import DeviceExperiences
    from pfrl_b200.utils.phi import Identity

    exp = np.load(os.path.join(GOLD, "ref_per_lazyframes_expected.npz"))
    per = PrioritizedReplayBuffer(capacity=40, alpha=0.6, beta0=0.4, betasteps=100)
    per.load(os.path.join(GOLD, "ref_per_lazyframes.pkl"))
    n = int(exp["n"])
    assert len(per) == n
    assert np.array_equal(per.store.read_priorities(), exp["priority"])
    info = per.store.info()
    assert info["max_priority"] == float(exp["max_priority"])       # b2rl_per_set_max_priority
    assert abs(info["total"] - float(exp["total"])) < 1e-9
    idx = torch.arange(n, dtype=torch.int64, device=per.device)
    b = batch_experiences(DeviceExperiences(per, n, index=idx), per.device, Identity(), 0.99)
    np.testing.assert_array_equal(b["state"].float().cpu().numpy(), exp["state"])
    np.testing.assert_array_equal(b["next_state"].float().cpu().numpy(), exp["next_state"])
    np.testing.assert_allclose(b["reward"].cpu().numpy(), exp["reward"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(b["discount"].cpu().numpy(), exp["discount"])
    assert np.array_equal(b["is_state_terminal"].cpu().numpy(), exp["is_state_terminal"])
    # a new transition is appended with the restored max_priority
    from pfrl_b200.utils.lazy_frames import LazyFrames

    frames = [np.full((1, 6, 6), i, np.uint8) for i in range(5)]
    per.append(LazyFrames(frames[:4], stack_axis=0), 1, 0.5, LazyFrames(frames[1:], stack_axis=0))
    per._flush()
    assert per.store.read_priorities()[-1] == float(exp["max_priority"])

    exp = np.load(os.path.join(GOLD, "ref_uniform_3step_expected.npz"))
    uni = ReplayBuffer(capacity=50, num_steps=3)
    uni.load(os.path.join(GOLD, "ref_uniform_3step.pkl"))
    n = int(exp["n"])
    idx = torch.arange(n, dtype=torch.int64, device=uni.device)
    b = batch_experiences(DeviceExperiences(uni, n, index=idx), uni.device, Identity(), 0.9)
    np.testing.assert_array_equal(b["state"].cpu().numpy(), exp["state"])
    np.testing.assert_array_equal(b["next_state"].cpu().numpy(), exp["next_state"])
    np.testing.assert_allclose(b["reward"].cpu().numpy(), exp["reward"], rtol=1e-6, atol=1e-7)


def test_a2c_on_cuda_matches_reference_trace():
    from test_a2c_cpu import G, N, _model
    from pfrl_b200.agents import A2C

    model = _model("discrete")
    opt = torch.optim.RMSprop(model.parameters(), lr=7e-3, eps=1e-5, alpha=0.99)
    agent = A2C(model, opt, gamma=0.97, num_processes=N, update_steps=4, gpu=0,
                average_actor_loss_decay=0.0, average_entropy_decay=0.0,
                average_value_decay=0.0, use_gae=False, max_grad_norm=0.5)
    # the CUDA sampler draws a different stream than the CPU one: feed the reference's actions
    steps = G["reward"].shape[0]
    for t in range(steps):
        agent.batch_act(list(G["obs"][t]))
        slot = agent.t - agent.t_start
        agent.window.actions[slot] = torch.tensor(G["discrete_actions"][t], dtype=torch.float32,
                                                 device=agent.device)
        agent.batch_observe(list(G["obs"][t + 1]), list(G["reward"][t]), list(G["done"][t]),
                            [False] * N)
        got = [v for _, v in agent.get_statistics()]
        np.testing.assert_allclose(got, G["discrete_stats"][t], rtol=1e-4, atol=1e-5)


def test_pinned_slab_uploads_in_one_copy():
    from test_vector_envs_cpu import WalkEnv
    from pfrl_b200.envs import MultiprocessVectorEnv
    from pfrl_b200.utils.batch_states import batch_states
    from pfrl_b200.utils.phi import Identity

    vec = MultiprocessVectorEnv([(lambda: WalkEnv()) for _ in range(4)], pin=True)
    try:
        obs = vec.reset()
        assert vec._pinned and obs.host_batch is not None
        dev = torch.device("cuda", 0)
        b = batch_states(obs, dev, Identity())
        assert b.is_cuda and tuple(b.shape) == (4, 6)
        np.testing.assert_array_equal(b.cpu().numpy(), np.stack(obs))
        obs2, _, _, _ = vec.step([0, 1, 2, 0])
        b2 = batch_states(obs2, dev, Identity())
        np.testing.assert_array_equal(b2.cpu().numpy(), np.stack(obs2))
        np.testing.assert_array_equal(b.cpu().numpy(), np.stack(obs))   # the first upload is a copy
    finally:
        vec.close()


def test_collections_prioritized_buffer_on_cuda():
    from test_collections_cpu import replay_collections_trace
    from pfrl_b200.collections import PrioritizedBuffer

    buf = replay_collections_trace(lambda cap: PrioritizedBuffer(capacity=cap))
    assert len(buf) == 300
