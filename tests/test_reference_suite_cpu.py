"""The reference's replay-buffer unit tests as re-expressed in
tests/test_reference_suite_gpu.py, run a second time WITHOUT a GPU: the CUDA
store is replaced by tests/fake_store.OracleBackedStore, so what is exercised
here is the buffers' host logic (windows, terminal tails, capacity, env_id
separation, save / load, the dict view, weight normalisation)."""
import os
import sys
from unittest import mock

import pytest

sys.path.insert(0, os.path.dirname(__file__))
import test_reference_suite_gpu as suite  # noqa: E402
from fake_store import OracleBackedStore  # noqa: E402


@pytest.fixture(autouse=True)
def _host_store():
    with mock.patch("pfrl_b200.replay_buffers.device_buffer.DeviceReplayStore", OracleBackedStore), \
            mock.patch("torch.cuda.current_device", return_value=0):
        yield


# collected from THIS module, so the gpu marker of the original module does not apply
TestReplayBuffer = suite.TestReplayBuffer
test_capacity_drops_the_oldest = suite.test_capacity_drops_the_oldest
test_env_id_windows_do_not_mix = suite.test_env_id_windows_do_not_mix
test_normalize_by_max = suite.test_normalize_by_max

# the buffer-level tests of tests/test_replay_buffers_gpu.py that do not create CUDA tensors
import test_replay_buffers_gpu as bufsuite  # noqa: E402

test_protocol_asserts_like_reference = bufsuite.test_protocol_asserts_like_reference
test_device_per_arbitrary_host_phi_and_dict_view = \
    bufsuite.test_device_per_arbitrary_host_phi_and_dict_view
test_save_load_round_trip = bufsuite.test_save_load_round_trip
test_lazyframes_buffer_saves_frames_once_and_accepts_appends_after_load = \
    bufsuite.test_lazyframes_buffer_saves_frames_once_and_accepts_appends_after_load
