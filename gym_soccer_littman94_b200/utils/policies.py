"""Table policies in the reference's format (gym_soccer/utils/policies.py:4-27): a dict
observation index -> action.  The same dicts are what SoccerVecEnv / SoccerSimultaneousEnv accept
as player_a_policy / player_b_policy and what the K2 kernel takes as on-device int8[nS] tables."""
import pickle

import numpy as np

NOOP = 0


def get_random_policy(n_states=761, n_actions=5, seed=0):
    """Same draws as the reference: one RandomState(seed).randint per state, in state order."""
    rs = np.random.RandomState(seed)
    return {s: rs.randint(0, n_actions) for s in range(n_states)}


def get_stand_policy(n_states=761):
    return {s: NOOP for s in range(n_states)}


def policy_to_table(policy, n_states=None):
    """dict / sequence -> contiguous int8 array (the device format)."""
    if isinstance(policy, dict):
        n_states = len(policy) if n_states is None else n_states
        return np.array([int(policy[s]) for s in range(n_states)], dtype=np.int8)
    return np.ascontiguousarray(policy, dtype=np.int8)


def save_policy(policy, filename, mode='wb'):
    assert isinstance(policy, dict), "Policy must be a dictionary"
    with open(filename, mode) as f:
        pickle.dump(policy, f)


def load_policy(filename, mode='rb'):
    with open(filename, mode) as f:
        return pickle.load(f)
