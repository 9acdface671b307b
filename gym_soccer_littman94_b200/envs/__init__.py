from .soccer_simultaneous_env import SoccerSimultaneousEnv  # noqa: F401  (reference: gym_soccer/envs/__init__.py:1)
from .vec_env import SoccerVecEnv  # noqa: F401
