from pfrl_b200.envs.serial_vector_env import SerialVectorEnv  # NOQA
from pfrl_b200.envs.synthetic import DeviceObsList  # NOQA
from pfrl_b200.envs.synthetic import SyntheticAtariVectorEnv  # NOQA
from pfrl_b200.envs.synthetic import SyntheticContinuousVectorEnv  # NOQA
from pfrl_b200.envs.toy import ChainEnv  # NOQA
from pfrl_b200.envs.multiprocess_vector_env import HostObsList, MultiprocessVectorEnv  # NOQA
from pfrl_b200.envs.abc import ABC  # NOQA
