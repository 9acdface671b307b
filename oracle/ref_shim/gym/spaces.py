"""gym.spaces.Dict / Discrete stand-ins: only `.n`, item access and `.sample()`."""
import numpy as np


class Discrete:
    def __init__(self, n, seed=None):
        self.n = int(n)
        self._rng = np.random.RandomState(seed)

    def sample(self):
        return int(self._rng.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    __contains__ = contains

    def __repr__(self):
        return f"Discrete({self.n})"


class Dict:
    def __init__(self, spaces=None, **kw):
        self.spaces = dict(spaces or {}, **kw)

    def __getitem__(self, k):
        return self.spaces[k]

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def items(self):
        return self.spaces.items()

    def values(self):
        return self.spaces.values()

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __repr__(self):
        return f"Dict({self.spaces})"
