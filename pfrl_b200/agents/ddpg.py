"""Module path of the reference (pfrl/agents/ddpg.py); the class lives next to
TD3, with which it shares the replay / gather / target-update machinery."""
from pfrl_b200.agents.td3 import DDPG  # NOQA
