"""ctypes binding of the CPU ORACLE (oracle/soccer_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT: importable only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  Nothing under
gym_soccer_littman94_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsoccer_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("soccer_oracle.c", "soccer_oracle.h")]
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(_SO) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsoccer_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


class _State(C.Structure):
    _fields_ = [("xa", C.c_int), ("ya", C.c_int), ("xb", C.c_int), ("yb", C.c_int), ("p", C.c_int)]

    def tup(self):
        return (self.xa, self.ya, self.xb, self.yb, self.p)


class _Trans(C.Structure):
    _fields_ = [("prob", C.c_double), ("ns", _State), ("obs", C.c_int),
                ("reward", C.c_double), ("done", C.c_int)]


class _Env(C.Structure):
    _fields_ = [("m", C.c_void_p), ("state", _State), ("timestep", C.c_int), ("needs_reset", C.c_int)]


STATE_DTYPE = np.dtype([("xa", "<i4"), ("ya", "<i4"), ("xb", "<i4"), ("yb", "<i4"), ("p", "<i4")])

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, i, d, i64, u64 = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_uint64
    ip = C.POINTER(C.c_int)
    L.orc_model_new.restype = vp
    L.orc_model_new.argtypes = [i, i, d, ip, ip]
    L.orc_model_free.argtypes = [vp]
    for name in ("orc_nS", "orc_nA", "orc_width", "orc_height", "orc_multiagent", "orc_n_goal_rows",
                 "orc_n_unreachable", "orc_n_goal_states", "orc_isd_len"):
        getattr(L, name).restype = i
        getattr(L, name).argtypes = [vp]
    L.orc_goal_row.restype = i
    L.orc_goal_row.argtypes = [vp, i]
    L.orc_isd_prob.restype = d
    L.orc_isd_prob.argtypes = [vp, i]
    L.orc_isd_state.restype = _State
    L.orc_isd_state.argtypes = [vp, i]
    L.orc_state_to_obs.restype = i
    L.orc_state_to_obs.argtypes = [vp, _State]
    L.orc_obs_to_state.restype = _State
    L.orc_obs_to_state.argtypes = [vp, i]
    L.orc_goal_reward.restype = d
    L.orc_goal_reward.argtypes = [vp, _State]
    L.orc_is_goal_state.restype = i
    L.orc_is_goal_state.argtypes = [vp, _State]
    L.orc_next_cell.argtypes = [vp, i, i, i, i, i, ip, ip]
    L.orc_transitions.restype = i
    L.orc_transitions.argtypes = [vp, _State, i, C.POINTER(C.POINTER(_Trans))]
    L.orc_dump_table.restype = i
    L.orc_dump_table.argtypes = [vp, i, vp, vp, vp, vp, vp, vp]
    L.orc_categorical_sample.restype = i
    L.orc_categorical_sample.argtypes = [C.POINTER(d), i, d]
    L.orc_fill_pmat_rmat.argtypes = [vp, vp, vp]
    L.orc_env_init.argtypes = [C.POINTER(_Env), vp]
    L.orc_env_reset.restype = i
    L.orc_env_reset.argtypes = [C.POINTER(_Env), d, C.POINTER(d)]
    L.orc_env_step.restype = i
    L.orc_env_step.argtypes = [C.POINTER(_Env), i, d, ip, C.POINTER(d), ip, ip, C.POINTER(d)]
    L.orc_rollout_injected.argtypes = [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i]
    L.orc_philox4x32_10.argtypes = [vp, vp, vp]
    L.orc_philox_word.restype = C.c_uint32
    L.orc_philox_word.argtypes = [u64, u64, u64]
    L.orc_philox_decode.argtypes = [C.c_uint32, ip, ip, ip, ip]
    L.orc_philox_r32.restype = C.c_uint32
    L.orc_philox_r32.argtypes = [C.c_uint32]
    L.orc_rollout_philox.argtypes = [vp, i64, i64, vp, vp, vp, vp, u64, u64, u64, vp, vp, vp, vp, i]
    _lib = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def n_states(width: int, height: int) -> int:
    f = width * height
    return 1 + 2 * f * (f - 1)


class OracleModel:
    """The reference's constructor products (SIM:35-144): state space, isd, transition table."""

    def __init__(self, width=5, height=4, slip_prob=0.0, player_a_policy=None, player_b_policy=None):
        L = lib()
        self._L = L

        def pol(p):
            if p is None:
                return None
            n = n_states(width, height)
            arr = (C.c_int * n)(*[int(p[s]) for s in range(n)])
            return arr
        pa, pb = pol(player_a_policy), pol(player_b_policy)
        self._m = L.orc_model_new(int(width), int(height), float(slip_prob),
                                  C.cast(pa, C.POINTER(C.c_int)) if pa is not None else None,
                                  C.cast(pb, C.POINTER(C.c_int)) if pb is not None else None)
        if not self._m:
            raise AssertionError("orc_model_new rejected the arguments (SIM:38,45-46)")
        self.nS = L.orc_nS(self._m)
        self.nA = L.orc_nA(self._m)
        self.width = L.orc_width(self._m)
        self.height = L.orc_height(self._m)
        self.multiagent = bool(L.orc_multiagent(self._m))
        self.single_key_is_b = player_a_policy is not None
        self.goal_rows = tuple(L.orc_goal_row(self._m, k) for k in range(L.orc_n_goal_rows(self._m)))
        self.goal_cols = (0, self.width - 1)
        self.n_unreachable = L.orc_n_unreachable(self._m)
        self.n_goal_states = L.orc_n_goal_states(self._m)
        self.isd = [(L.orc_isd_prob(self._m, k), L.orc_isd_state(self._m, k).tup())
                    for k in range(L.orc_isd_len(self._m))]
        self.nkeys = 25 if self.multiagent else 5

    def __del__(self):
        try:
            if getattr(self, "_m", None):
                self._L.orc_model_free(self._m)
                self._m = None
        except Exception:
            pass

    def state_to_obs(self, st):
        return self._L.orc_state_to_obs(self._m, _State(*st))

    def obs_to_state(self, obs):
        return self._L.orc_obs_to_state(self._m, int(obs)).tup()

    def is_goal_state(self, st):
        return bool(self._L.orc_is_goal_state(self._m, _State(*st)))

    def goal_reward(self, st):
        return self._L.orc_goal_reward(self._m, _State(*st))

    def next_cell(self, x, y, move, has_ball):
        nx, ny = C.c_int(), C.c_int()
        self._L.orc_next_cell(self._m, x, y, move[0], move[1], int(bool(has_ball)), C.byref(nx), C.byref(ny))
        return nx.value, ny.value

    def transitions(self, st, key):
        """P_readable[st][key] as a list of (prob, next_tuple, reward, done, obs)."""
        p = C.POINTER(_Trans)()
        n = self._L.orc_transitions(self._m, _State(*st), int(key), C.byref(p))
        if n < 0:
            raise KeyError((st, key))
        return [(p[k].prob, p[k].ns.tup(), p[k].reward, bool(p[k].done), p[k].obs) for k in range(n)]

    def dump_table(self):
        """Whole P / P_readable table as padded arrays [nS, nkeys, L], list order preserved."""
        L = self._L.orc_dump_table(self._m, 0, None, None, None, None, None, None)
        nS, nk = self.nS, self.nkeys
        d = dict(count=np.zeros((nS, nk), np.uint8), prob=np.zeros((nS, nk, L), np.float64),
                 next_obs=np.zeros((nS, nk, L), np.int32), reward=np.zeros((nS, nk, L), np.int8),
                 done=np.zeros((nS, nk, L), np.uint8), next_tuple=np.full((nS, nk, L, 5), -1, np.int8))
        rc = self._L.orc_dump_table(self._m, L, _ptr(d["count"]), _ptr(d["prob"]), _ptr(d["next_obs"]),
                                    _ptr(d["reward"]), _ptr(d["done"]), _ptr(d["next_tuple"]))
        assert rc == L
        return d

    def pmat_rmat(self):
        shp = (self.nA, self.nA) if self.multiagent else (self.nA,)
        P = np.empty((self.nS, self.nS) + shp, dtype=np.float64)
        R = np.empty((self.nS,) + shp, dtype=np.float64)
        self._L.orc_fill_pmat_rmat(self._m, _ptr(P), _ptr(R))
        return P, R

    def states_from_obs(self, obs):
        out = np.zeros(len(obs), dtype=STATE_DTYPE)
        for k, o in enumerate(obs):
            out[k] = self.obs_to_state(int(o))
        return out

    def obs_from_states(self, states):
        return np.array([self.state_to_obs(tuple(int(v) for v in s)) for s in states], dtype=np.int32)

    def rollout_injected(self, states, timesteps, act_a, act_b, rng8, rng32=None, n_threads=1,
                         want_reset_obs=True, out=None):
        """Lock-step auto-reset rollout.  act/rng arrays are [T, N]; states/timesteps [N] are updated in place.
        out = (obs, rew, flg, ro) reuses result arrays (bench.py's CPU timing: no page faults in the timed loop)."""
        T, N = act_a.shape
        assert states.dtype == STATE_DTYPE and timesteps.dtype == np.int32
        for a in (act_a, act_b, rng8):
            assert a is None or (a.dtype == np.uint8 and a.flags.c_contiguous and a.shape == (T, N))
        assert rng32 is None or (rng32.dtype == np.uint32 and rng32.shape == (T, N))
        if out is not None:
            obs, rew, flg, ro = out
            assert obs.shape == (T, N) and obs.dtype == np.int32 and rew.dtype == np.float32 and flg.dtype == np.uint8
        else:
            obs = np.empty((T, N), np.int32)
            rew = np.empty((T, N), np.float32)
            flg = np.empty((T, N), np.uint8)
            ro = np.empty((T, N), np.int32) if want_reset_obs else None
        self._L.orc_rollout_injected(self._m, T, N, _ptr(states), _ptr(timesteps), _ptr(act_a), _ptr(act_b),
                                     _ptr(rng8), _ptr(rng32), _ptr(obs), _ptr(rew), _ptr(flg), _ptr(ro),
                                     int(n_threads))
        return obs, rew, flg, ro

    def rollout_philox(self, states, timesteps, K, seed, step0=0, env_id_base=0, policy_a=None,
                       policy_b=None, n_threads=1, want_streams=True):
        N = len(states)
        assert states.dtype == STATE_DTYPE and timesteps.dtype == np.int32
        obs = np.empty((K, N), np.int32) if want_streams else None
        rew = np.empty((K, N), np.float32) if want_streams else None
        flg = np.empty((K, N), np.uint8) if want_streams else None
        stats = np.zeros(6, np.int64)
        pa = None if policy_a is None else np.ascontiguousarray(policy_a, dtype=np.int8)
        pb = None if policy_b is None else np.ascontiguousarray(policy_b, dtype=np.int8)
        self._L.orc_rollout_philox(self._m, K, N, _ptr(states), _ptr(timesteps), _ptr(pa), _ptr(pb),
                                   int(seed), int(step0), int(env_id_base),
                                   _ptr(obs), _ptr(rew), _ptr(flg), _ptr(stats), int(n_threads))
        return obs, rew, flg, stats


class OracleEnv:
    """One mutable env over an OracleModel: reset(u) / step(key, u) with the draw supplied."""

    def __init__(self, model: OracleModel):
        self.model = model
        self._e = _Env()
        lib().orc_env_init(C.byref(self._e), model._m)

    @property
    def state(self):
        return self._e.state.tup()

    @state.setter
    def state(self, st):
        self._e.state = _State(*st)

    @property
    def timestep(self):
        return self._e.timestep

    @timestep.setter
    def timestep(self, v):
        self._e.timestep = int(v)

    @property
    def needs_reset(self):
        return bool(self._e.needs_reset)

    @needs_reset.setter
    def needs_reset(self, v):
        self._e.needs_reset = int(bool(v))

    def reset(self, u):
        p = C.c_double()
        obs = lib().orc_env_reset(C.byref(self._e), float(u), C.byref(p))
        return obs, p.value

    def step(self, key, u):
        o, d, t = C.c_int(), C.c_int(), C.c_int()
        r, p = C.c_double(), C.c_double()
        rc = lib().orc_env_step(C.byref(self._e), int(key), float(u), C.byref(o), C.byref(r),
                                C.byref(d), C.byref(t), C.byref(p))
        if rc == -1:
            raise AssertionError("Please reset the environment before taking a step")
        if rc != 0:
            raise KeyError(rc)
        return o.value, r.value, bool(d.value), bool(t.value), p.value


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def philox_word(seed, env_id, step):
    return int(lib().orc_philox_word(int(seed), int(env_id), int(step)))


def philox_r32(w):
    """The 32-bit step draw of a word: lo32(25 w); u = (r32 + 0.5) / 2^32."""
    return int(lib().orc_philox_r32(int(w)))


def philox_decode(w):
    a, b, s, r = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    lib().orc_philox_decode(int(w), C.byref(a), C.byref(b), C.byref(s), C.byref(r))
    return a.value, b.value, s.value, r.value
