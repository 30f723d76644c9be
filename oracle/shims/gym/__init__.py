"""oracle shim: gym.spaces.Box only (common_agent.py:3, hrl_agent.py:3)."""
from . import spaces
