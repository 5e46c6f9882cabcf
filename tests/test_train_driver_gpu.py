"""GPU: train_agent_batch's episode accounting on the device (EpisodeLedger) gives the
numbers of the host accounting for the same reward / done stream, and the driver keeps
the reference's call-count contract with a device-resident vector env
(pfrl/experiments/train_agent_batch.py:65-141)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_device_ledger_equals_host_ledger():
    from pfrl_b200.experiments.train_agent_batch import EpisodeLedger

    rng = np.random.RandomState(0)
    E = 7
    host, dev = EpisodeLedger(E, 5), EpisodeLedger(E, 5)
    for t in range(300):
        r = rng.randn(E)
        d = rng.rand(E) < 0.08
        infos = [{"needs_reset": bool(rng.rand() < 0.03)} for _ in range(E)]
        a = host.advance(r, d, infos, 40)
        b = dev.advance(torch.as_tensor(r, device="cuda"), torch.as_tensor(d, device="cuda"),
                        infos, 40)
        assert dev.on_device and not host.on_device
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))
        assert host.episodes == dev.episodes
        np.testing.assert_array_equal(host.recent_returns(), dev.recent_returns())
        assert np.array_equal(host.restart_finished(), dev.restart_finished())


def test_driver_contract_with_device_env(tmp_path):
    from unittest import mock

    from pfrl_b200.experiments import train_agent_batch

    E = 4

    class DevEnv:
        num_envs = E

        def __init__(self):
            self.t = 0
            self.resets = []

        def reset(self, mask=None):
            self.resets.append(None if mask is None else np.asarray(mask).copy())
            return [torch.zeros(3, device="cuda") for _ in range(E)]

        def step(self, actions):
            self.t += 1
            done = torch.tensor([self.t % 5 == 0, False, self.t % 3 == 0, False], device="cuda")
            return ([torch.full((3,), float(self.t), device="cuda") for _ in range(E)],
                    torch.full((E,), 0.5, device="cuda"), done, [{} for _ in range(E)])

        def close(self):
            pass

    agent = mock.Mock()
    agent.batch_act.return_value = [0] * E
    agent.get_statistics.return_value = []
    hook = mock.Mock()
    env = DevEnv()
    train_agent_batch(agent, env, steps=40, outdir=str(tmp_path), step_hooks=[hook],
                      max_episode_len=7)
    assert agent.batch_act.call_count == 10 and agent.batch_observe.call_count == 10
    assert [c[0][2] for c in hook.call_args_list] == list(range(1, 41))
    # done flags reach the agent as host bools, reset masks follow done / max_episode_len
    _, r, d, rs = agent.batch_observe.call_args_list[4][0]     # vector step 5
    assert list(d) == [True, False, False, False] and r.dtype == np.float64
    assert list(env.resets[5]) == [False, True, True, True]    # env 0 done at t=5
    # env 1 and 3 never report done: they are reset by max_episode_len = 7
    assert list(agent.batch_observe.call_args_list[6][0][3]) == [False, True, False, True]
    agent.save.assert_called_once()
