"""gym_soccer_littman94_b200 -- the Littman'94 soccer step/reset hot path on B200.

Drop-in for the hot path of mimoralea/gym-soccer-littman94:
    from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv   # single env, same API
    from gym_soccer_littman94_b200.envs import SoccerVecEnv            # batched, new
    env = gym_soccer_littman94_b200.make("SoccerSimultaneous-v0")      # the id the reference leaves
                                                                        # commented out
"""

__all__ = ["make", "REGISTRY"]

# The registration block the reference keeps commented out (gym_soccer/__init__.py:5-12).
REGISTRY = {
    "SoccerSimultaneous-v0": dict(
        entry_point="gym_soccer_littman94_b200.envs:SoccerSimultaneousEnv",
        kwargs={"width": 5, "height": 4, "slip_prob": 0.2, "player_a_policy": None, "player_b_policy": None},
        max_episode_steps=100, reward_threshold=1.0, nondeterministic=True),
}


def make(id="SoccerSimultaneous-v0", num_envs=None, **kwargs):
    """gym.make-style factory.  num_envs=None -> the single-env drop-in; an int -> SoccerVecEnv."""
    if id not in REGISTRY:
        raise KeyError(f"unknown environment id {id!r}; known: {sorted(REGISTRY)}")
    from .envs import SoccerSimultaneousEnv, SoccerVecEnv
    kw = dict(REGISTRY[id]["kwargs"])
    kw.update(kwargs)
    if num_envs is None:
        return SoccerSimultaneousEnv(**kw)
    return SoccerVecEnv(int(num_envs), **kw)
