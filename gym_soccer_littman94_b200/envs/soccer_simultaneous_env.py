"""SoccerSimultaneousEnv -- single-env drop-in for the reference class of the same name
(SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py), executing every
transition on the GPU through the C ABI in include/soccer_b200.h.

Surface kept identical to the reference:
  constructor kwargs and asserts (SIM:35-46); attributes width (padded), height, slip_prob,
  nS, nA, multiagent, return_agent, goal_rows, goal_cols, state_space, goal_states,
  unreachable_states, isd, P, P_readable, Pmat, Rmat, observation_space, action_space,
  np_random and the writable state / timestep / needs_reset / lastaction / observations;
  reset(seed, options) -> (obs, infos) (SIM:410-424); step(action) -> (obs, rewards, dones,
  truncateds, infos) with the reference's assert messages (SIM:375-408); render (SIM:426-485);
  _state_to_observation / _observation_to_state (SIM:487-497); the action constants (SIM:8-32).

Randomness: exactly one `self.np_random.random()` draw per step() and per reset(), like the
reference (SIM:395, 414), so with the same seed the two produce the same trajectory.  The
draw is handed to the kernel: as its 2-bit quantisation floor(4u) when slip_prob == 0 (exact:
outcome lists have 1, 2 or 4 equiprobable entries) and as the raw fp64 otherwise.

P / P_readable / Pmat / Rmat are built lazily from the sweep and dense kernels (K3).
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import weakref

import numpy as np
import torch

from .. import _lib, spaces
from .._lib import Pitch, StepArgs, check


def _drain(stream, *buffers):
    """Finalizer of an env: wait for its last (speculative) launch; `buffers` are held until then."""
    try:
        stream.synchronize()
    except Exception:  # noqa: BLE001  (interpreter shutdown: the context may be gone)
        pass


_REC = struct.Struct('<IifI')       # one record of soccer_step_speculate: state word, obs, reward, detail flags


class SoccerSimultaneousEnv:
    # SIM:8-32
    NOOP = 0
    NORTH = 1
    SOUTH = 2
    EAST = 3
    WEST = 4
    ACTION_STRING = ['NOOP', 'NORTH', 'SOUTH', 'EAST', 'WEST']
    ACTION_STRING_TO_INT = {k: v for v, k in enumerate(ACTION_STRING)}
    ACTION_INT_TO_MOVE = {NOOP: (0, 0), NORTH: (0, -1), SOUTH: (0, 1), EAST: (1, 0), WEST: (-1, 0)}
    ACTION_STRING_TO_MOVE = {'NOOP': (0, 0), 'NORTH': (0, -1), 'SOUTH': (0, 1), 'EAST': (1, 0), 'WEST': (-1, 0)}
    MOVE_TO_ACTION_STRING = {v: k for k, v in ACTION_STRING_TO_MOVE.items()}
    MOVE_TO_ACTION_INT = {v: k for k, v in ACTION_INT_TO_MOVE.items()}
    TERMINAL_STATE = (-1, -1, -1, -1, -1)

    # byte offsets inside the 32-byte host<->device mailbox
    _OFF_U, _OFF_ACT_A, _OFF_ACT_B, _OFF_RNG8 = 0, 8, 9, 10
    _OFF_OBS, _OFF_REWARD, _OFF_FLAGS, _OFF_STATE = 16, 20, 24, 28
    _BUF_BYTES = 32
    _STATE_MASK = 0x0100FFFF      # cells + possession; timestep / needs_reset live on the host

    def __init__(self, width=5, height=4, slip_prob=0.0, player_a_policy=None, player_b_policy=None, seed=0,
                 device="cuda"):
        assert not (player_a_policy is not None and player_b_policy is not None), \
            "Both players cannot have a policy. At least one must be None."
        assert width >= 5, "Width must be at least 5 columns."
        assert height >= 4, "Height must be at least 4 rows."

        self._lib = _lib.lib()
        self._pitch = Pitch(int(width), int(height), float(slip_prob))
        self._info = _lib.pitch_info(width, height, slip_prob)
        self._field_width = int(width)
        self.width = width + 2
        self.height = height
        self.slip_prob = slip_prob
        self.seed = seed
        self.player_a_policy = player_a_policy
        self.player_b_policy = player_b_policy
        self.multiagent = player_a_policy is None and player_b_policy is None
        self.return_agent = ['player_a', 'player_b'] if self.multiagent else ['player_a'] \
            if player_a_policy is None else ['player_b']
        self._n_agents = len(self.return_agent)
        self.np_random = np.random.RandomState()
        self.np_random.seed(self.seed)

        self.goal_rows = tuple(self._info.goal_rows[i] for i in range(self._info.n_goal_rows))
        self.goal_cols = (0, self.width - 1)
        self._enumerate_states()
        self.nA = len(self.ACTION_STRING)
        self.observation_space = spaces.Dict({a: spaces.Discrete(self.nS) for a in self.return_agent})
        self.action_space = spaces.Dict({a: spaces.Discrete(self.nA) for a in self.return_agent})
        self.isd = [(1.0 / self._info.n_isd, tuple(self._info.isd_tuple[i][k] for k in range(5)))
                    for i in range(self._info.n_isd)]
        self._mp = [self._info.slip_combo_prob[c] for c in range(9)]

        # device side: a 32-byte mailbox (inputs at 0..15, outputs + the state word at 16..31)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SoccerB200Error("SoccerSimultaneousEnv needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._dbuf = torch.zeros(self._BUF_BYTES, dtype=torch.uint8, device=self.device)
        self._hbuf = torch.zeros(self._BUF_BYTES, dtype=torch.uint8).pin_memory()
        self._hnp = self._hbuf.numpy()
        self._staged = os.environ.get("SOCCER_B200_SINGLE_ENV_STAGED", "0") == "1"
        # typed views of the mailbox fields and per-call objects built once: step() is latency-bound
        # (one launch + one synchronize), so the host side is kept to a handful of scalar stores
        h = self._hnp
        self._v_u = h[self._OFF_U:self._OFF_U + 8].view(np.float64)
        self._v_state = h[self._OFF_STATE:self._OFF_STATE + 4].view(np.uint32)
        self._v_obs = h[self._OFF_OBS:self._OFF_OBS + 4].view(np.int32)
        self._v_reward = h[self._OFF_REWARD:self._OFF_REWARD + 4].view(np.float32)
        self._step_args = None
        self._round_cache = {}
        self._pol_a = self._policy_tensor(player_a_policy)
        self._pol_b = self._policy_tensor(player_b_policy)
        self._stream = torch.cuda.Stream(device=self.device)
        self._stream_ptr = C.c_void_p(self._stream.cuda_stream)
        # speculative step (slip_prob == 0, soccer_step_speculate): as soon as the state is known ONE launch steps it for
        # all 25 joint actions x 4 draw values into this pinned buffer; step() then only picks a record.
        # SOCCER_B200_SINGLE_ENV_SPECULATE=0 keeps every step on the launch-and-wait path (A/B, tests).
        # slip_prob > 0: the draw is a full fp64 number and cannot be enumerated, so the launch is handed the draw the
        # env's generator is GOING to make: a shadow RandomState with the same state runs one draw ahead (_draw()), and
        # every draw of np_random is checked against it -- a generator the caller reseeded, replaced or drew from
        # shows up as a mismatch, the speculation is dropped (launch-and-wait step) and the shadow resynchronised.
        self._spec_on = (not self._staged and os.environ.get("SOCCER_B200_SINGLE_ENV_SPECULATE", "1") == "1")
        self._shadow, self._u_ahead, self._spec_u, self._c_u = None, None, None, C.c_double(0.0)
        self._c_u_ref = C.byref(self._c_u)
        self._shadow_of = None
        if self._spec_on and slip_prob != 0:
            self._shadow = np.random.RandomState()
            self._resync()
        # sequence numbers start at a per-env random value: a record left in a recycled pinned block by another env's
        # launch cannot pass for one of this env's
        self._spec_key, self._spec_seq = None, int.from_bytes(os.urandom(3), "little")
        self._p_ref = C.byref(self._pitch)
        if self._spec_on:
            self._sbuf = torch.zeros(1600, dtype=torch.uint8).pin_memory()
            self._s_bytes = memoryview(self._sbuf.numpy())          # records: struct '<IifI' at 16 * index
            self._s_u32 = self._s_bytes.cast('I')                  # word 3 of a record: detail flags | seq << 8
            self._s_ptr = C.c_void_p(self._sbuf.data_ptr())
            self._s_pol = (C.c_void_p(self._pol_a.data_ptr()) if self._pol_a is not None else None,
                           C.c_void_p(self._pol_b.data_ptr()) if self._pol_b is not None else None)
        torch.cuda.synchronize(self.device)      # mailbox zero-fill and policy uploads done before the first launch
        # the kernels write into the pinned buffers behind torch's back (raw launches): drain the env's stream before
        # the buffers can go back to the caching host allocator
        weakref.finalize(self, _drain, self._stream, self._hbuf, self._dbuf, getattr(self, "_sbuf", None))

        self._tables = None     # lazily built (P, P_readable)
        self._dense = None      # lazily built (Pmat, Rmat)
        self.needs_reset = True
        self._state_word = None
        self.timestep = 0
        self.observations = None
        self.lastaction = None

    # ------------------------------------------------------------------ constructor products
    def _enumerate_states(self):
        """SIM:63-109: same nested order and the same three classes; the index of a reachable
        tuple is the closed form the kernels use."""
        H, W, w = self.height, self.width, self._field_width
        F = w * H
        gr = set(self.goal_rows)
        self.unreachable_states, self.goal_states = [], {}
        self.state_space = {self.TERMINAL_STATE: 0}
        for xa in range(H):
            for ya in range(W):
                a_goal_col = ya in self.goal_cols
                for xb in range(H):
                    for yb in range(W):
                        b_goal_col = yb in self.goal_cols
                        for p in range(2):
                            st = (xa, ya, xb, yb, p)
                            if (a_goal_col and xa not in gr) or (b_goal_col and xb not in gr):
                                self.unreachable_states.append(st)
                            elif (a_goal_col and p != 0) or (b_goal_col and p != 1):
                                self.unreachable_states.append(st)
                            elif xa == xb and ya == yb:
                                self.unreachable_states.append(st)
                            elif a_goal_col or b_goal_col:
                                holder_col = ya if p == 0 else yb
                                self.goal_states[st] = 1.0 if holder_col == W - 1 else -1.0
                            else:
                                a, b = xa * w + ya - 1, xb * w + yb - 1
                                self.state_space[st] = 1 + 2 * (a * (F - 1) + b - (1 if b > a else 0)) + p
        self.nS = len(self.state_space)
        assert self.nS == self._info.nS
        self._reverse_state_space = {v: k for k, v in self.state_space.items()}

    def _policy_tensor(self, policy):
        if policy is None:
            return None
        arr = np.array([int(policy[s]) for s in range(self.nS)], dtype=np.int8)
        return torch.from_numpy(arr).to(self.device)

    # ------------------------------------------------------------------ state attribute
    @property
    def state(self):
        if self._state_word is None:
            return None
        tup = (C.c_int32 * 5)()
        check(self._lib.soccer_unpack_state_host(C.byref(self._pitch), C.c_uint32(self._state_word), C.byref(tup),
                                                 None, None), "soccer_unpack_state_host")
        return tuple(int(v) for v in tup)

    @state.setter
    def state(self, st):
        if st is None:
            self._state_word = None
            return
        tup = (C.c_int32 * 5)(*[int(v) for v in st])
        word = C.c_uint32()
        rc = self._lib.soccer_pack_state_host(C.byref(self._pitch), C.byref(tup), 0, 0, C.byref(word))
        if rc != 0:
            raise KeyError(tuple(st))        # the reference fails with KeyError at SIM:394
        self._state_word = int(word.value) & self._STATE_MASK

    # ------------------------------------------------------------------ device round trip
    def _roundtrip(self, launch):
        """One kernel launch per call.  The 32-byte mailbox is pinned host memory, which under
        unified addressing the kernel reads and writes directly over PCIe (zero copy): one launch +
        one stream synchronize instead of copy / launch / copy / synchronize.
        SOCCER_B200_SINGLE_ENV_STAGED=1 restores the staged copies (A/B measurements)."""
        dev = self.device
        if torch.cuda.current_device() != dev.index:
            with torch.cuda.device(dev):
                return self._roundtrip(launch)
        cur = self._stream          # the env's own stream: everything it touches lives in the mailbox
        if self._staged:
            with torch.cuda.stream(cur):
                self._dbuf.copy_(self._hbuf, non_blocking=True)
                launch(self._stream_ptr, self._dbuf.data_ptr())
                self._hbuf.copy_(self._dbuf, non_blocking=True)
        else:
            launch(self._stream_ptr, self._hbuf.data_ptr())
        cur.synchronize()

    def _resync(self):
        """Give the shadow generator np_random's state (None: np_random is not a RandomState any more -> no speculation)."""
        self._u_ahead = None
        try:
            self._shadow.set_state(self.np_random.get_state())
            self._shadow_of = self.np_random
        except Exception:  # noqa: BLE001
            self._shadow = None

    def _draw(self):
        """The ONE np_random.random() of SIM:395 / SIM:414, with the shadow kept in step.  Returns (u, in_sync)."""
        u = self.np_random.random()
        sh = self._shadow
        if sh is None:
            return u, False
        ua = self._u_ahead
        if ua is None:
            ua = sh.random()
        self._u_ahead = None
        if ua != u or self._shadow_of is not self.np_random:
            self._resync()
            return u, False
        return u, True

    def _speculate(self):
        """Enqueue the step of the CURRENT state for every (joint action, draw) -- returns at once."""
        self._spec_key = None
        if not self._spec_on or self.needs_reset or self._state_word is None or not (0 <= self.timestep < 100):
            return
        word = (self._state_word & self._STATE_MASK) | (int(self.timestep) << 16)
        u_ref = None
        if self.slip_prob != 0:
            if self._shadow is None or self._u_ahead is not None:
                return
            self._u_ahead = self._spec_u = self._c_u.value = self._shadow.random()    # the draw the next step() will make
            u_ref = self._c_u_ref
        self._spec_seq = seq = (self._spec_seq + 1) & 0xFFFFFF or 1
        spec = self._lib.soccer_step_speculate
        rc = spec(self._p_ref, word, self._s_pol[0], self._s_pol[1], u_ref, self._s_ptr, seq, self._stream_ptr)
        if rc == 400:               # cudaErrorInvalidResourceHandle: another device is current -- retry under a guard
            with torch.cuda.device(self.device):
                rc = spec(self._p_ref, word, self._s_pol[0], self._s_pol[1], u_ref, self._s_ptr, seq, self._stream_ptr)
        if rc == 0:                 # (a goal tuple injected through `env.state = ...` is refused: launch-and-wait path)
            self._spec_key = word

    def _speculated(self, word, idx):
        """Record `idx` of the speculation launched for state `word`, or None if there is none."""
        if self._spec_key != word or self._spec_key is None:
            return None
        flag, seq, w = self._s_u32, self._spec_seq, idx * 4 + 3
        spins = 0
        while flag[w] >> 8 != seq:                    # usually already there: the launch was enqueued a call ago
            spins += 1
            if spins > 20000:
                self._stream.synchronize()
                if flag[w] >> 8 != seq:
                    return None
        # the sequence number travelled in the same 16-byte write as the record: read the record only now
        new_word, obs, reward, flags = _REC.unpack_from(self._s_bytes, idx * 16)
        return new_word, obs, reward, flags & 0xFF

    def reset(self, seed=None, options=None):
        if seed is not None:
            self.np_random.seed(seed)
            if self._shadow is not None:
                self._resync()
        u, _ = self._draw()                                          # the one draw of SIM:414
        n_isd = self._info.n_isd
        idx = min(int(u * n_isd), n_isd - 1)                         # argmax(cumsum > u), exact for 4 / 2 equal parts
        h = self._hnp
        h[self._OFF_RNG8] = ((idx if n_isd == 4 else idx << 1) & 3) << 2

        def launch(stream, base):
            check(self._lib.soccer_reset(C.byref(self._pitch), C.c_void_p(base + self._OFF_STATE),
                                         C.c_void_p(base + self._OFF_OBS), C.c_void_p(base + self._OFF_RNG8),
                                         None, 1, stream), "soccer_reset")
        self._roundtrip(launch)
        self._state_word = int(h[self._OFF_STATE:self._OFF_STATE + 4].view(np.uint32)[0]) & self._STATE_MASK
        obs = int(h[self._OFF_OBS:self._OFF_OBS + 4].view(np.int32)[0])
        p = self.isd[idx][0]
        self.observations = {a: obs for a in self.return_agent}
        infos = {a: {"p": np.round(p, 2)} for a in self.return_agent}
        self.lastaction = None
        self.needs_reset = False
        self.timestep = 0
        self._speculate()
        return self.observations, infos

    def _check_action(self, action):
        """The assert ladder of SIM:376-391, message for message."""
        assert isinstance(action, dict), "Action must be a dictionary"
        assert len(action) == 1 or len(action) == 2, "Action must be a dictionary of length 1 or 2"
        assert self.multiagent or self.player_a_policy is not None or self.player_b_policy is not None, \
            "Multiagent environment or policy for one player must be provided"
        assert self.player_a_policy is not None or 'player_a' in action, "A policy for player_a must be provided"
        assert self.player_b_policy is not None or 'player_b' in action, "A policy for player_b must be provided"
        if self.multiagent:
            assert (isinstance(action, dict) and len(action) == 2), \
                "Action must be a dictionary of length 2 for multiagent case"
            assert 'player_a' in action and 'player_b' in action, "Action must contain both 'player_a' and 'player_b'"
        else:
            assert (isinstance(action, dict) and len(action) == 1), \
                "Action must be a dictionary of length 1 for single agent case"
            assert 'player_a' in action or 'player_b' in action, "Action must contain either 'player_a' or 'player_b'"
            assert not ('player_a' in action and 'player_b' in action), \
                "Action must contain only one of 'player_a' or 'player_b'"

    def step(self, action):
        assert not self.needs_reset, "Please reset the environment before taking a step"
        # a well-formed action (the common case) passes every assert of SIM:376-391: skip the ladder
        if not (type(action) is dict and len(action) == self._n_agents and self.return_agent[0] in action
                and self.return_agent[-1] in action):
            self._check_action(action)
        aa = 0 if self._pol_a is not None else int(action['player_a'])
        ab = 0 if self._pol_b is not None else int(action['player_b'])
        if not (0 <= aa < self.nA and 0 <= ab < self.nA):
            raise IndexError("list index out of range")               # ACTION_STRING[...] at SIM:393
        if self._state_word is None:
            raise KeyError(None)

        u, in_sync = self._draw()                                    # the one draw of SIM:395
        h = self._hnp
        r2 = min(int(u * 4.0), 3)                                    # floor(4u): exact 2-bit form of u
        word = (self._state_word & self._STATE_MASK) | ((int(self.timestep) & 0xFF) << 16)
        spec = None
        if self._spec_key is not None:
            if self.slip_prob == 0:
                spec = self._speculated(word, (aa * 5 + ab) * 4 + r2)
            elif in_sync and u == self._spec_u:                      # the launch was handed exactly this draw
                spec = self._speculated(word, (aa * 5 + ab) * 4)
        if spec is not None:
            new_word, obs, reward, flags = spec
            return self._finish_step(action, new_word, obs, reward, flags)
        self._v_u[0] = u
        self._v_state[0] = word
        h[self._OFF_ACT_A], h[self._OFF_ACT_B] = aa, ab
        h[self._OFF_RNG8] = r2

        if self._step_args is None:
            base = self._dbuf.data_ptr() if self._staged else self._hbuf.data_ptr()
            a = StepArgs()
            a.state = base + self._OFF_STATE
            a.act_a = None if self._pol_a is not None else base + self._OFF_ACT_A
            a.act_b = None if self._pol_b is not None else base + self._OFF_ACT_B
            a.rng8 = base + self._OFF_RNG8
            a.rngf64 = base + self._OFF_U if self.slip_prob != 0 else None
            a.policy_a = None if self._pol_a is None else self._pol_a.data_ptr()
            a.policy_b = None if self._pol_b is None else self._pol_b.data_ptr()
            a.obs, a.reward, a.flags = base + self._OFF_OBS, base + self._OFF_REWARD, base + self._OFF_FLAGS
            a.reset_obs = None
            a.n, a.auto_reset, a.use_philox, a.detail = 1, 0, 0, 1
            self._step_args = (a, C.byref(a), C.byref(self._pitch))
        _, a_ref, p_ref = self._step_args

        def launch(stream, base):
            check(self._lib.soccer_step_ex(p_ref, a_ref, stream), "soccer_step_ex")
        self._roundtrip(launch)

        return self._finish_step(action, int(self._v_state[0]), int(self._v_obs[0]), float(self._v_reward[0]),
                                 int(h[self._OFF_FLAGS]))

    def _finish_step(self, action, new_word, obs, reward, flags):
        """The bookkeeping and dict-shaped returns of SIM:396-408 from one step record of the device."""
        done = bool(flags & 1)
        self._state_word = new_word & self._STATE_MASK
        self.timestep += 1
        trunc = self.timestep >= 100
        self.needs_reset = done or trunc
        self._speculate()           # the next step's launch goes out before the dicts below are built
        prob = self._mp[(flags >> 4) & 0xF] * (1.0, 0.5, 0.25)[(flags >> 2) & 3]   # mp * nsp, SIM:241
        p2 = self._round_cache.get(prob)
        if p2 is None:
            p2 = self._round_cache[prob] = np.round(prob, 2)         # SIM:405; a handful of distinct values
        self.lastaction = action
        if self.multiagent:                                          # literal dicts: same keys, same order, less interpreter time
            self.observations = {'player_a': obs, 'player_b': obs}
            return (self.observations, {'player_a': reward, 'player_b': reward * -1}, {'player_a': done, 'player_b': done},
                    {'player_a': trunc, 'player_b': trunc}, {'player_a': {"p": p2}, 'player_b': {"p": p2}})
        a = self.return_agent[0]
        self.observations = {a: obs}
        return self.observations, {a: reward}, {a: done}, {a: trunc}, {a: {"p": p2}}

    # ------------------------------------------------------------------ transition tables (K3)
    def _sweep(self):
        """Run the sweep kernel; returns numpy arrays shaped [nS-1, 25, C, ...]."""
        nC = 1 if self.slip_prob == 0 else 9
        n = (self.nS - 1) * 25 * nC
        dev = self.device
        n_out = torch.empty(n, dtype=torch.uint8, device=dev)
        nstate = torch.empty(n * 4, dtype=torch.int32, device=dev)
        nobs = torch.empty(n * 4, dtype=torch.int32, device=dev)
        rew = torch.empty(n * 4, dtype=torch.int8, device=dev)
        done = torch.empty(n * 4, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(self._lib.soccer_sweep(C.byref(self._pitch), nC, C.c_void_p(n_out.data_ptr()),
                                         C.c_void_p(nstate.data_ptr()), C.c_void_p(nobs.data_ptr()),
                                         C.c_void_p(rew.data_ptr()), C.c_void_p(done.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "soccer_sweep")
        shp = (self.nS - 1, 25, nC)
        return (n_out.cpu().numpy().reshape(shp), nstate.cpu().numpy().view(np.uint32).reshape(shp + (4,)),
                nobs.cpu().numpy().reshape(shp + (4,)), rew.cpu().numpy().reshape(shp + (4,)),
                done.cpu().numpy().reshape(shp + (4,)))

    def _word_to_tuple(self, word):
        w = self._field_width

        def dec(c):
            if c & 0x80:
                return (c & 0x3F, w + 1 if c & 0x40 else 0)
            return (c // w, c % w + 1)
        xa, ya = dec(word & 0xFF)
        xb, yb = dec((word >> 8) & 0xFF)
        return (xa, ya, xb, yb, (word >> 24) & 1)

    def _build_tables(self):
        """P and P_readable (SIM:167-293) from the sweep kernel's output, list order preserved."""
        n_out, nstate, nobs, rew, done = self._sweep()
        nC = n_out.shape[2]
        combos = [c for c in range(9) if self._mp[c] != 0] if nC == 9 else [0]     # SIM:226-227
        nsp = {1: 1.0, 2: 0.5, 4: 0.25}
        flip = (not self.multiagent) and ('player_b' in self.return_agent)          # SIM:243-244
        names = self.ACTION_STRING
        P, P_readable = {}, {}
        pol_a, pol_b = self.player_a_policy, self.player_b_policy
        for st, s in self.state_space.items():
            if s == 0:
                continue
            P[s], P_readable[st] = {}, {}
            aaa = range(self.nA) if pol_a is None else [pol_a[s]]
            aab = range(self.nA) if pol_b is None else [pol_b[s]]
            for aa in aaa:
                for ab in aab:
                    ja = aa * 5 + ab
                    tl, tlr = [], []
                    for c in combos:
                        k_n = int(n_out[s - 1, ja, c])
                        pr = self._mp[c] * nsp[k_n]
                        for k in range(k_n):
                            r = float(rew[s - 1, ja, c, k])
                            if flip:
                                r = -1 * r
                            d = bool(done[s - 1, ja, c, k])
                            tl.append((pr, int(nobs[s - 1, ja, c, k]), r, d))
                            tlr.append((pr, self._word_to_tuple(int(nstate[s - 1, ja, c, k])), r, d))
                    if self.multiagent:
                        P[s][(aa, ab)] = tl
                        P_readable[st][(names[aa], names[ab])] = tlr
                    elif pol_b is not None:
                        P[s][aa] = tl
                        P_readable[st][names[aa]] = tlr
                    else:
                        P[s][ab] = tl
                        P_readable[st][names[ab]] = tlr
        # goal states (SIM:235-236, 300-301): absorbing self-loops, one entry per non-zero slip
        # combination; P[0] holds the entry the last-enumerated goal state wrote (SIM:182-183)
        zero = -0.0 if flip else 0.0
        for st in self.goal_states:
            P[0], P_readable[st] = {}, {}
            aaa = range(self.nA) if pol_a is None else [pol_a[0]]
            aab = range(self.nA) if pol_b is None else [pol_b[0]]
            for aa in aaa:
                for ab in aab:
                    tl = [(self._mp[c] * 1.0, 0, zero, True) for c in range(9) if self._mp[c] != 0]
                    tlr = [(self._mp[c] * 1.0, st, zero, True) for c in range(9) if self._mp[c] != 0]
                    if self.multiagent:
                        P[0][(aa, ab)] = tl
                        P_readable[st][(names[aa], names[ab])] = tlr
                    elif pol_b is not None:
                        P[0][aa] = tl
                        P_readable[st][names[aa]] = tlr
                    else:
                        P[0][ab] = tl
                        P_readable[st][names[ab]] = tlr
        self._tables = (P, P_readable)

    def _build_dense(self):
        """Pmat / Rmat (SIM:170-171, 258-279) from the dense kernel, fp64, reference order."""
        dev = self.device
        shp = (self.nA, self.nA) if self.multiagent else (self.nA,)
        Pm = torch.empty((self.nS, self.nS) + shp, dtype=torch.float64, device=dev)
        Rm = torch.empty((self.nS,) + shp, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(self._lib.soccer_dense(
                C.byref(self._pitch), None if self._pol_a is None else C.c_void_p(self._pol_a.data_ptr()),
                None if self._pol_b is None else C.c_void_p(self._pol_b.data_ptr()), C.c_void_p(Pm.data_ptr()),
                C.c_void_p(Rm.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "soccer_dense")
        self._dense = (Pm.cpu().numpy(), Rm.cpu().numpy())

    @property
    def P(self):
        if self._tables is None:
            self._build_tables()
        return self._tables[0]

    @property
    def P_readable(self):
        if self._tables is None:
            self._build_tables()
        return self._tables[1]

    @property
    def Pmat(self):
        if self._dense is None:
            self._build_dense()
        return self._dense[0]

    @property
    def Rmat(self):
        if self._dense is None:
            self._build_dense()
        return self._dense[1]

    # ------------------------------------------------------------------ index maps, SIM:487-497
    def _state_to_observation(self, state):
        state = self.TERMINAL_STATE if state in self.goal_states else state
        return self.state_space[state]

    def _observation_to_state(self, observation):
        return self._reverse_state_space[observation]

    # ------------------------------------------------------------------ render, SIM:426-485
    def render(self):
        st = self.state
        print(st)
        xa, ya, xb, yb, p = st
        print(f"Player A position: x={xa}, y={ya}, possession={p == 0}")
        print(f"Player B position: x={xb}, y={yb}, possession={p == 1}")
        grid = [[' '] * self.width for _ in range(self.height)]
        grid[xa][ya] = 'A*' if p == 0 else 'A '
        grid[xb][yb] = 'B*' if p == 1 else 'B '
        bar = '  ' + '-' * (self.width * 2 - 4)
        lines = [bar]
        for ri, row in enumerate(grid):
            cells = [f'{c:<2}' for c in row]
            if ri not in self.goal_rows:
                lines.append(' |' + ''.join(cells[1:-1]) + '| ')
            elif '*' in row[0]:
                lines.append(''.join(cells[:-1]) + '||')
            elif '*' in row[-1]:
                lines.append('||' + ''.join(cells[1:]))
            else:
                lines.append('||' + ''.join(cells[1:-1]) + '||')
        lines.append(bar)
        for ln in lines:
            print(ln)
        print(f"Ball possession: {'A' if p == 0 else 'B'}")
        if self.lastaction and self.multiagent:
            action_a, action_b = self.lastaction.values()
            print(f"Last actions: A: {self.ACTION_STRING[action_a]}, B: {self.ACTION_STRING[action_b]}")
        elif self.lastaction:
            who = 'A' if self.player_a_policy is None else 'B'
            act = self.lastaction['player_a' if who == 'A' else 'player_b']
            print(f"Last action: {who}: {self.ACTION_STRING[act]}")
        holder_x, holder_y = (xa, ya) if p == 0 else (xb, yb)
        if holder_x in self.goal_rows and holder_y in self.goal_cols:
            scored_right = holder_y == self.width - 1
            if p == 0:
                print("GOAL! Player A scored!" if scored_right else "OWN GOAL! Player A scored in their own goal!")
            else:
                print("OWN GOAL! Player B scored in their own goal!" if scored_right else "GOAL! Player B scored!")
