"""`spaces.Dict` / `spaces.Discrete` for observation_space / action_space (SIM:126-131).

Uses gym or gymnasium when one is installed (so isinstance checks in user code hold);
otherwise a minimal local stand-in with the attributes the reference's tests touch
(`.n`, item access, `in`, `.sample()`).
"""
try:  # pragma: no cover - depends on the environment
    from gym.spaces import Dict, Discrete  # type: ignore
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gymnasium.spaces import Dict, Discrete  # type: ignore
    except Exception:  # noqa: BLE001
        import numpy as _np

        class Discrete:  # type: ignore
            def __init__(self, n, seed=None):
                self.n = int(n)
                self._rng = _np.random.RandomState(seed)

            def sample(self):
                return int(self._rng.randint(self.n))

            def contains(self, x):
                return 0 <= int(x) < self.n

            __contains__ = contains

            def __repr__(self):
                return f"Discrete({self.n})"

        class Dict:  # type: ignore
            def __init__(self, spaces=None, **kw):
                self.spaces = dict(spaces or {}, **kw)

            def __getitem__(self, k):
                return self.spaces[k]

            def __contains__(self, k):
                return k in self.spaces

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
