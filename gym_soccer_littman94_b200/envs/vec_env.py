"""SoccerVecEnv -- N lock-step Littman'94 soccer environments on one B200.

The batched, drop-in counterpart of the reference's single-env path
(SIM = /root/reference/gym_soccer/envs/soccer_simultaneous_env.py): same observation index
(SIM:487-497), same action encoding (SIM:8-12), same rules (SIM:296-373), same reward / done /
truncation (SIM:235-240, 399-406); `reset` (SIM:410-424) is fused into `step` (auto-reset).

PyTorch supplies device memory and streams only; every transition is computed by the CUDA
kernels behind include/soccer_b200.h.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from .. import _lib
from .._lib import Pitch, StepArgs, check

LAYOUT_CELL, LAYOUT_INDEX = 0, 1
HOST_WIDE, HOST_NARROW, HOST_PACKED = 0, 1, 2     # soccer_step_host_args.narrow
INITIAL_RESET_STEP = (1 << 64) - 1   # Philox step index reserved for the very first reset draw


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class SoccerVecEnv:
    """num_envs independent soccer games stepped by one kernel launch.

    Parameters mirror the reference constructor (SIM:35) plus the batch controls:
      num_envs, device
      rng_mode   "injected": the caller passes the draws (bit-exact replay of the reference);
                 "philox":   Philox4x32-10 keyed (seed, env_id_base + i, step)
      kernel     "rules": rules evaluated inline (any pitch / option);
                 "table": transition table resident in shared memory (5x4 and 6x4 pitches; multi-agent or
                          single-agent with the folded player's table policy next to the table);
                 "auto":  table when it applies, else rules
      env_id_base  global id of env 0 (rank * envs_per_rank when sharded over GPUs; a multiple of 4 keeps the
                   Philox kernels on their 4-envs-per-thread path, see include/soccer_b200.h)

    Pitch limit of this library (the reference has none): width * height <= 126 and height <= 16 -- one-byte cell
    codes and 16-bit observation lanes in the kernels; larger pitches raise ValueError.
    """

    def __init__(self, num_envs: int, width: int = 5, height: int = 4, slip_prob: float = 0.0,
                 player_a_policy=None, player_b_policy=None, seed: int = 0,
                 device="cuda", rng_mode: str = "injected", kernel: str = "auto",
                 env_id_base: int = 0, want_reset_obs: bool = True):
        assert not (player_a_policy is not None and player_b_policy is not None), \
            "Both players cannot have a policy. At least one must be None."          # SIM:38
        assert width >= 5, "Width must be at least 5 columns."                       # SIM:45
        assert height >= 4, "Height must be at least 4 rows."                        # SIM:46
        assert rng_mode in ("injected", "philox") and kernel in ("auto", "rules", "table")
        if int(width) * int(height) > 126 or int(height) > 16:
            raise ValueError(f"pitch {width}x{height} is beyond this library's limit (width * height <= 126, height <= 16: "
                             "one-byte cell codes, 16-bit observation lanes); the reference itself has no upper limit")
        self.lib = _lib.lib()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.SoccerB200Error("SoccerVecEnv needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.pitch = Pitch(int(width), int(height), float(slip_prob))
        self.info = _lib.pitch_info(width, height, slip_prob)
        self.width = self.info.padded_width                                          # SIM:48
        self.height = int(height)
        self.slip_prob = float(slip_prob)
        self.nS, self.nA = self.info.nS, self.info.nA
        self.seed = int(seed)
        self.rng_mode = rng_mode
        self.env_id_base = int(env_id_base)
        self.multiagent = player_a_policy is None and player_b_policy is None        # SIM:54
        self.return_agent = ['player_a', 'player_b'] if self.multiagent else \
            (['player_a'] if player_a_policy is None else ['player_b'])              # SIM:55-56
        self.policy_a = self._policy_tensor(player_a_policy)
        self.policy_b = self._policy_tensor(player_b_policy)
        self.want_reset_obs = bool(want_reset_obs)
        self._pitch_ref = C.byref(self.pitch)
        self._plain = self.rng_mode == "injected" and self.multiagent and self.slip_prob == 0.0
        self._fast_args = None

        _nb = C.c_int64()
        table_ok = self.lib.soccer_step_table_bytes_host(C.byref(self.pitch), C.byref(_nb)) == 0
        if table_ok and not self.multiagent:                 # the folded policies ride next to the table (6x4: no room)
            table_ok = _nb.value + 16 + 2 * ((self.nS + 15) & ~15) <= 227 * 1024 - 1024
        if kernel == "table" and not table_ok:
            raise _lib.SoccerB200Error("kernel='table' needs a table that fits shared memory (5x4 and 6x4 pitches; "
                                       "5x4 with a folded policy)")
        # "auto": the table kernel pays a 152 KB shared-memory fill per CTA per launch, which only
        # amortises over large batches; small batches are launch-latency bound either way
        if kernel == "auto":
            # (with slip_prob > 0 the rules walk costs ~500 instructions per env, the table walk ~120: switch early)
            big = self.num_envs >= (65536 if self.slip_prob == 0.0 else 4096)
            kernel = "table" if (table_ok and big) else "rules"
        self.kernel = kernel
        self.layout = LAYOUT_INDEX if self.kernel == "table" else LAYOUT_CELL

        n, dev = self.num_envs, self.device
        self.state = torch.zeros(n, dtype=torch.int32, device=dev)       # packed uint32 words
        self.obs = torch.zeros(n, dtype=torch.int32, device=dev)
        self.reward = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flags = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.reset_obs = torch.zeros(n, dtype=torch.int32, device=dev) if self.want_reset_obs else None
        self.step_count = 0
        self.table = None
        self.slip_index = None
        if self.kernel == "table":
            nbytes = C.c_int64()
            check(self.lib.soccer_step_table_bytes_host(C.byref(self.pitch), C.byref(nbytes)), "step_table_bytes")
            self.table = torch.zeros(nbytes.value // 2, dtype=torch.int16, device=dev)
            with torch.cuda.device(dev):
                check(self.lib.soccer_build_step_table(C.byref(self.pitch), _ptr(self.table), _stream(dev)),
                      "soccer_build_step_table")
                if self.slip_prob != 0.0:
                    # slip index: first slip combination with more than one outcome per (obs, joint action); lets
                    # soccer_step_table_slip decide most draws from constant prefix sums
                    ib = C.c_int64()
                    check(self.lib.soccer_slip_index_bytes_host(C.byref(self.pitch), C.byref(ib)), "slip_index_bytes")
                    self.slip_index = torch.zeros(ib.value, dtype=torch.uint8, device=dev)
                    check(self.lib.soccer_build_slip_index(C.byref(self.pitch), _ptr(self.table), _ptr(self.slip_index),
                                                           _stream(dev)), "soccer_build_slip_index")
                # the step kernels prefetch the table BEFORE their grid-dependency wait (programmatic
                # dependent launch), so it must be complete before the first step is enqueued
                torch.cuda.current_stream(dev).synchronize()

    # ------------------------------------------------------------------ helpers
    def _policy_tensor(self, policy):
        if policy is None:
            return None
        if isinstance(policy, dict):                         # utils/policies.py:4-15 format
            policy = [int(policy[s]) for s in range(self.nS)]
        t = torch.as_tensor(policy, dtype=torch.int8).reshape(-1)
        assert t.numel() == self.nS, "a table policy needs one action per observation index"
        return t.to(self.device).contiguous()

    def _check_vec(self, t, dtype, name, host=False):
        """A [num_envs] stream the kernels read or write through its raw pointer: dtype, device, contiguity and
        length are checked here because the C ABI cannot (a strided row or a wrong dtype would be an out-of-bounds
        device access).  host=True: a PINNED CPU tensor instead (zero copy: the kernel moves it over PCIe)."""
        if t is None:
            return None
        ok = isinstance(t, torch.Tensor) and t.dtype == dtype and t.is_contiguous() and t.numel() == self.num_envs
        if ok and host:
            ok = (not t.is_cuda) and t.is_pinned()
        elif ok:
            ok = t.is_cuda and t.device == self.device
        if not ok:
            where = "pinned CPU" if host else f"CUDA ({self.device})"
            raise ValueError(f"{name} must be a contiguous {dtype} {where} tensor with {self.num_envs} elements")
        return t

    def _check_out(self, out, host=False):
        obs, reward, flags, reset_obs = out
        return (self._check_vec(obs, torch.int32, "out obs", host), self._check_vec(reward, torch.float32, "out reward", host),
                self._check_vec(flags, torch.uint8, "out flags", host), self._check_vec(reset_obs, torch.int32, "out reset_obs", host))

    def _to_layout(self):
        if self.layout == LAYOUT_INDEX:
            check(self.lib.soccer_convert_state(C.byref(self.pitch), _ptr(self.state), _ptr(self.state),
                                                LAYOUT_INDEX, self.num_envs, _stream(self.device)),
                  "soccer_convert_state")

    # ------------------------------------------------------------------ API
    def reset(self, rng8: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """reset() of every env (SIM:410-424).  injected mode: start state from rng8 bits 2..3."""
        if self.num_envs == 0:
            return self.obs
        with torch.cuda.device(self.device):
            st = _stream(self.device)
            cell_state = self.state
            if self.layout == LAYOUT_INDEX and mask is not None:
                raise NotImplementedError("masked reset is only offered on the rules kernel")
            if self.rng_mode == "injected":
                rng8 = self._check_vec(rng8, torch.uint8, "rng8")
                if rng8 is None:
                    raise ValueError("rng_mode='injected' needs rng8")
                check(self.lib.soccer_reset(C.byref(self.pitch), _ptr(cell_state), _ptr(self.obs), _ptr(rng8),
                                            _ptr(mask), self.num_envs, st), "soccer_reset")
            else:
                check(self.lib.soccer_reset_philox(C.byref(self.pitch), _ptr(cell_state), _ptr(self.obs),
                                                   _ptr(mask), self.seed, INITIAL_RESET_STEP, self.env_id_base,
                                                   self.num_envs, st), "soccer_reset_philox")
            self._to_layout()
        self.step_count = 0
        return self.obs

    def set_state(self, obs: torch.Tensor, timestep: Optional[torch.Tensor] = None):
        """`env.state = ...` for every env, from observation indices (SIM:496-497)."""
        obs = self._check_vec(obs, torch.int32, "obs")
        timestep = self._check_vec(timestep, torch.int32, "timestep")
        with torch.cuda.device(self.device):
            check(self.lib.soccer_set_state(C.byref(self.pitch), _ptr(self.state), _ptr(obs), _ptr(timestep),
                                            self.num_envs, _stream(self.device)), "soccer_set_state")
            self._to_layout()

    def current_obs(self) -> torch.Tensor:
        """_state_to_observation of every env's current state (SIM:487-494)."""
        if self.layout == LAYOUT_INDEX:
            return (self.state & 0xFFFF).to(torch.int32)
        out = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.soccer_get_obs(C.byref(self.pitch), _ptr(self.state), _ptr(out), self.num_envs,
                                          _stream(self.device)), "soccer_get_obs")
        return out

    def timesteps(self) -> torch.Tensor:
        return (self.state >> 16) & 0xFF

    def step(self, act_a: Optional[torch.Tensor], act_b: Optional[torch.Tensor] = None,
             rng8: Optional[torch.Tensor] = None, rng32: Optional[torch.Tensor] = None,
             rngf64: Optional[torch.Tensor] = None, out=None, detail: bool = False, auto_reset: bool = True,
             stats: Optional[torch.Tensor] = None, _host: bool = False):
        """One lock-step step() of all envs (SIM:375-408) + fused reset (SIM:410-424).

        Returns (obs, reward, flags, reset_obs) device tensors: obs is what the reference's
        step() returned (0 on a goal), reward the first return agent's reward, flags bit 0
        terminated / bit 1 truncated, reset_obs the observation the next step starts from.
        `out` may supply the four output tensors (e.g. rows of a [T, N] buffer).
        `stats` (int64[6] CUDA tensor): accumulate this step's episode statistics inside the step kernel
        ([episodes, goals_A, goals_B, truncations, steps, sum_episode_len], the vector rollout() returns).
        """
        if out is not None:
            obs, reward, flags, reset_obs = self._check_out(out, _host)
        else:
            obs, reward, flags, reset_obs = self.obs, self.reward, self.flags, self.reset_obs
        if self.num_envs == 0:
            return obs, reward, flags, reset_obs
        # fast host path for the plain lock-step step (injected draws, slip 0, no folded policy, no options): small
        # batches are bound by this Python code, so it makes one ctypes call with integer pointers and nothing else
        if self._plain and auto_reset and not detail and rng32 is None and rngf64 is None \
                and act_a is not None and act_b is not None and rng8 is not None \
                and torch.cuda.current_device() == self.device.index:
            u8 = torch.uint8
            cv = lambda t, d, nm: self._check_vec(t, d, nm, _host)      # noqa: E731
            st = torch.cuda.current_stream(self.device).cuda_stream
            ro = None if reset_obs is None else reset_obs.data_ptr()
            if stats is not None:
                # the same step with the statistics fused into the kernel: one soccer_step_ex call on a cached argument
                # block (a step of 2^24 envs takes 51 us on the GPU: the host side must stay well below that)
                if not (stats.is_cuda and stats.dtype == torch.int64 and stats.is_contiguous() and stats.numel() >= 6):
                    raise ValueError("stats must be a contiguous int64 CUDA tensor with 6 elements")
                a = self._fast_args
                if a is None:
                    a = self._fast_args = StepArgs()
                    a.state, a.n, a.auto_reset = self.state.data_ptr(), self.num_envs, 1
                    a.table = None if self.table is None else self.table.data_ptr()
                    self._fast_ref = C.byref(a)
                a.act_a, a.act_b = cv(act_a, u8, "act_a").data_ptr(), cv(act_b, u8, "act_b").data_ptr()
                a.rng8 = cv(rng8, u8, "rng8").data_ptr()
                a.obs, a.reward, a.flags, a.reset_obs = obs.data_ptr(), reward.data_ptr(), flags.data_ptr(), ro
                a.stats = stats.data_ptr()
                rc = self.lib.soccer_step_ex(self._pitch_ref, self._fast_ref, st)
                if rc:
                    check(rc, "soccer_step_ex")
                self.step_count += 1
                return obs, reward, flags, reset_obs
            if self.kernel == "table":
                rc = self.lib.soccer_step_table(self._pitch_ref, self.table.data_ptr(), self.state.data_ptr(),
                                                cv(act_a, u8, "act_a").data_ptr(), cv(act_b, u8, "act_b").data_ptr(),
                                                cv(rng8, u8, "rng8").data_ptr(), obs.data_ptr(), reward.data_ptr(),
                                                flags.data_ptr(), ro, self.num_envs, st)
            else:
                rc = self.lib.soccer_step(self._pitch_ref, self.state.data_ptr(),
                                          cv(act_a, u8, "act_a").data_ptr(), cv(act_b, u8, "act_b").data_ptr(),
                                          cv(rng8, u8, "rng8").data_ptr(), obs.data_ptr(), reward.data_ptr(),
                                          flags.data_ptr(), ro, self.num_envs, st)
            if rc:
                check(rc, "soccer_step")
            self.step_count += 1
            return obs, reward, flags, reset_obs
        cv = lambda t, d, nm: self._check_vec(t, d, nm, _host)          # noqa: E731
        if stats is not None and not (isinstance(stats, torch.Tensor) and stats.is_cuda and stats.dtype == torch.int64
                                      and stats.is_contiguous() and stats.numel() >= 6):
            raise ValueError("stats must be a contiguous int64 CUDA tensor with 6 elements")
        with torch.cuda.device(self.device):
            st = _stream(self.device)
            philox = self.rng_mode == "philox"
            # the table kernels serve everything but the per-outcome detail flags and auto_reset=False (an INDEX-layout
            # state cannot hold needs_reset); fused statistics are not offered together with slip_prob > 0 there
            on_table = self.kernel == "table" and not detail and auto_reset
            fused_stats = stats is not None and not (on_table and self.slip_prob != 0.0)
            state = self.state
            if self.kernel == "table" and not on_table:
                if not auto_reset:
                    raise NotImplementedError("kernel='table' states cannot hold needs_reset; use kernel='rules' "
                                              "for auto_reset=False")
                # a table env keeps INDEX-layout states: the generic rules kernel runs on a CELL-layout copy
                state = torch.empty_like(self.state)
                check(self.lib.soccer_convert_state(C.byref(self.pitch), _ptr(self.state), _ptr(state), LAYOUT_CELL,
                                                    self.num_envs, st), "soccer_convert_state")
            a = StepArgs()
            a.state = state.data_ptr()
            a.act_a = None if self.policy_a is not None else cv(act_a, torch.uint8, "act_a").data_ptr()
            a.act_b = None if self.policy_b is not None else cv(act_b, torch.uint8, "act_b").data_ptr()
            if not philox:
                a.rng8 = cv(rng8, torch.uint8, "rng8").data_ptr()
                if self.slip_prob != 0.0 and rng32 is None and rngf64 is None:
                    raise ValueError("slip_prob > 0 needs the step draw: rng32 (int32/uint32 bits) or rngf64")
                if rngf64 is not None:
                    a.rngf64 = cv(rngf64, torch.float64, "rngf64").data_ptr()
                elif rng32 is not None:
                    a.rng32 = cv(rng32, torch.int32, "rng32").data_ptr()
            a.policy_a = None if self.policy_a is None else self.policy_a.data_ptr()
            a.policy_b = None if self.policy_b is None else self.policy_b.data_ptr()
            a.obs, a.reward, a.flags = obs.data_ptr(), reward.data_ptr(), flags.data_ptr()
            a.reset_obs = None if reset_obs is None else reset_obs.data_ptr()
            a.n = self.num_envs
            a.auto_reset = 1 if auto_reset else 0
            a.use_philox = 1 if philox else 0
            a.detail = 1 if detail else 0
            a.seed, a.step, a.env_id_base = self.seed, self.step_count, self.env_id_base
            a.stats = stats.data_ptr() if fused_stats else None
            if on_table:
                a.table = self.table.data_ptr()
                use_index = self.slip_index is not None and os.environ.get("SOCCER_B200_SLIP_WALK", "0") != "1"
                a.slip_index = self.slip_index.data_ptr() if use_index else None
            check(self.lib.soccer_step_ex(C.byref(self.pitch), C.byref(a), st), "soccer_step_ex")
            if state is not self.state:
                check(self.lib.soccer_convert_state(C.byref(self.pitch), _ptr(state), _ptr(self.state), LAYOUT_INDEX,
                                                    self.num_envs, st), "soccer_convert_state")
            if stats is not None and not fused_stats:
                # separate pass over the streams (counts goals by the sign of the STREAMED reward: swap them back for a
                # player_b env, whose reward stream is negated; sum_episode_len is not available here)
                tmp = torch.zeros(6, dtype=torch.int64, device=self.device)
                check(self.lib.soccer_step_stats(_ptr(flags), _ptr(reward), self.num_envs, _ptr(tmp), st), "soccer_step_stats")
                if self.policy_a is not None:
                    tmp = tmp[[0, 2, 1, 3, 4, 5]]
                stats[:6] += tmp
        self.step_count += 1
        return obs, reward, flags, reset_obs

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self) -> dict:
        """Everything a run needs to resume: the packed state words (CELL layout, whatever kernel the env uses) and the
        Philox coordinates.  Philox is counter-based -- a pure function of (seed, global env id, step) -- so a restored
        env continues the very trajectory, on any GPU count and with either kernel (the reference checkpoints nothing
        but policy pickles, utils/policies.py:17-27)."""
        state = self.state
        if self.layout == LAYOUT_INDEX and self.num_envs:
            state = torch.empty_like(self.state)
            with torch.cuda.device(self.device):
                check(self.lib.soccer_convert_state(C.byref(self.pitch), _ptr(self.state), _ptr(state), LAYOUT_CELL,
                                                    self.num_envs, _stream(self.device)), "soccer_convert_state")
        return {"state": state.detach().cpu().clone(), "layout": LAYOUT_CELL, "step_count": int(self.step_count),
                "seed": int(self.seed), "env_id_base": int(self.env_id_base), "rng_mode": self.rng_mode,
                "width": int(self.pitch.width), "height": int(self.pitch.height), "slip_prob": float(self.slip_prob),
                "num_envs": int(self.num_envs)}

    def load_state_dict(self, sd: dict) -> None:
        if (sd["width"], sd["height"], sd["num_envs"]) != (int(self.pitch.width), int(self.pitch.height), self.num_envs) \
                or float(sd["slip_prob"]) != self.slip_prob:
            raise ValueError("checkpoint belongs to a different pitch, slip_prob or batch size")
        state = sd["state"].to(self.device, dtype=torch.int32).contiguous()
        if self.layout == LAYOUT_INDEX and self.num_envs:
            with torch.cuda.device(self.device):
                check(self.lib.soccer_convert_state(C.byref(self.pitch), _ptr(state), _ptr(self.state), LAYOUT_INDEX,
                                                    self.num_envs, _stream(self.device)), "soccer_convert_state")
        else:
            self.state.copy_(state)
        self.step_count, self.seed, self.env_id_base = int(sd["step_count"]), int(sd["seed"]), int(sd["env_id_base"])

    # ------------------------------------------------------------------ the reference's dict-shaped surface, batched
    def reset_dict(self, seed=None, options=None, rng8: Optional[torch.Tensor] = None):
        """reset() with the reference's signature and return shape (SIM:410-424): (obs_dict, info_dict) keyed by
        agent, every value an [N] tensor.  `seed` reseeds like the reference's reset(seed) (Philox key)."""
        if seed is not None:
            self.seed = int(seed)
        obs = self.reset(rng8)
        p = torch.full((self.num_envs,), float(np.round(1.0 / self.info.n_isd, 2)), dtype=torch.float64,
                       device=self.device)
        return {a: obs for a in self.return_agent}, {a: {"p": p} for a in self.return_agent}

    def step_dict(self, action: dict, rng8: Optional[torch.Tensor] = None, rng32: Optional[torch.Tensor] = None):
        """step() with the reference's signature and return shape (SIM:375-408) for the whole batch:
        `action` = {'player_a': uint8[N], 'player_b': uint8[N]} (one key in the single-agent modes); returns
        (obs, rewards, dones, truncateds, infos) dicts keyed by agent whose values are [N] tensors -- obs int32
        (0 on a goal, like the reference), rewards float32 (B = -A in the multi-agent case, SIM:401-402), dones /
        truncateds bool, infos {"p": float64 probability of the realised outcome rounded to 2 places, SIM:405}
        (kernel="rules" only; the table kernel does not track it).  Episodes restart by themselves: the
        observation a finished env continues from is `self.reset_obs`."""
        assert isinstance(action, dict), "Action must be a dictionary"
        assert len(action) == 1 or len(action) == 2, "Action must be a dictionary of length 1 or 2"
        assert self.policy_a is not None or 'player_a' in action, "A policy for player_a must be provided"
        assert self.policy_b is not None or 'player_b' in action, "A policy for player_b must be provided"
        if self.multiagent:
            assert len(action) == 2, "Action must be a dictionary of length 2 for multiagent case"
        else:
            assert len(action) == 1, "Action must be a dictionary of length 1 for single agent case"
        detail = self.kernel == "rules"
        obs, reward, flags, _ = self.step(action.get('player_a'), action.get('player_b'), rng8=rng8, rng32=rng32,
                                          detail=detail)
        rewards = {a: reward for a in self.return_agent}
        if self.multiagent:
            rewards['player_b'] = -reward
        done, trunc = (flags & 1).bool(), (flags & 2).bool()
        infos = {a: {} for a in self.return_agent}
        if detail:
            if getattr(self, "_p_lut", None) is None:
                mp = [self.info.slip_combo_prob[c] for c in range(9)] + [0.0] * 7
                lut = [float(np.round(mp[i >> 2] * (1.0, 0.5, 0.25, 0.0)[i & 3], 2)) for i in range(64)]   # SIM:241, 405
                self._p_lut = torch.tensor(lut, dtype=torch.float64, device=self.device)
            p = self._p_lut[(flags >> 2).long()]
            infos = {a: {"p": p} for a in self.return_agent}
        return ({a: obs for a in self.return_agent}, rewards, {a: done for a in self.return_agent},
                {a: trunc for a in self.return_agent}, infos)

    def step_many(self, act_a: torch.Tensor, act_b: torch.Tensor, rng8: torch.Tensor, out=None):
        """T lock-steps from [T, N] uint8 CUDA tensors, enqueued by one C call (no per-step Python
        round trip; capturable into a CUDA graph).  Returns [T, N] obs / reward / flags / reset_obs."""
        if self.slip_prob != 0.0 or not self.multiagent or self.rng_mode != "injected":
            raise NotImplementedError("step_many covers the multi-agent, slip_prob == 0, injected-draw step")
        T, n, dev = act_a.shape[0], self.num_envs, self.device
        for name, t in (("act_a", act_a), ("act_b", act_b), ("rng8", rng8)):
            if not (t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous() and tuple(t.shape) == (T, n)):
                raise ValueError(f"{name} must be a contiguous uint8 CUDA tensor of shape ({T}, {n})")
        if out is None:
            out = (torch.empty((T, n), dtype=torch.int32, device=dev), torch.empty((T, n), dtype=torch.float32, device=dev),
                   torch.empty((T, n), dtype=torch.uint8, device=dev),
                   torch.empty((T, n), dtype=torch.int32, device=dev) if self.want_reset_obs else None)
        obs, reward, flags, reset_obs = out
        if n and T:
            with torch.cuda.device(dev):
                check(self.lib.soccer_step_many(C.byref(self.pitch), _ptr(self.table), _ptr(self.state), T, _ptr(act_a),
                                                _ptr(act_b), _ptr(rng8), _ptr(obs), _ptr(reward), _ptr(flags),
                                                _ptr(reset_obs), n, _stream(dev)), "soccer_step_many")
        self.step_count += T
        return obs, reward, flags, reset_obs

    def step_stats(self, flags: Optional[torch.Tensor] = None, reward: Optional[torch.Tensor] = None,
                   stats: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Accumulate [episodes, goals_A, goals_B, truncations, steps, -] of one step's streams."""
        flags = self.flags if flags is None else flags
        reward = self.reward if reward is None else reward
        if stats is None:
            stats = torch.zeros(6, dtype=torch.int64, device=self.device)
        if flags.numel() == 0:
            return stats
        with torch.cuda.device(self.device):
            check(self.lib.soccer_step_stats(_ptr(flags), _ptr(reward), flags.numel(), _ptr(stats), _stream(self.device)),
                  "soccer_step_stats")
        return stats

    # ------------------------------------------------------------------ host-buffer path (end to end)
    def _pinned(self, *dtypes):
        """num_envs-element pinned host tensors, one per dtype, carved from ONE huge-page-backed allocation
        (_lib.HostArena: full-rate PCIe reads, unlike small individual pin_memory() allocations)."""
        n = self.num_envs
        sizes = [n * torch.empty(0, dtype=d).element_size() for d in dtypes]
        arena = _lib.HostArena(sum((sz + 4095) // 4096 * 4096 for sz in sizes) + 4096)
        return tuple(arena.take(n, d) for d in dtypes)

    def alloc_host_inputs(self, packed: bool = False):
        """One set of pinned host input buffers for step_host / step_host_packed: (act_a, act_b, rng8), or
        (joint, rng8) with packed=True -- uint8[num_envs] CPU tensors for the caller to fill."""
        return self._pinned(*([torch.uint8] * (2 if packed else 3)))

    def _host_buffers(self, narrow: bool):
        key = "_hb_narrow" if narrow else "_hb_wide"
        if getattr(self, key, None) is None:
            n, dev = self.num_envs, self.device
            if getattr(self, "_host_common", None) is None:
                nbytes = C.c_int64()
                check(self.lib.soccer_step_host_scratch_bytes_host(n, C.byref(nbytes)), "scratch_bytes")
                self._host_common = dict(scratch=torch.empty(nbytes.value, dtype=torch.uint8, device=dev),
                                         streams=[torch.cuda.Stream(device=dev) for _ in range(3)],
                                         h_flags=self._pinned(torch.uint8)[0])
            h_obs, h_reward = self._pinned(*((torch.int16, torch.int8) if narrow else (torch.int32, torch.float32)))
            setattr(self, key, dict(h_obs=h_obs, h_reward=h_reward))
        return self._host_common, getattr(self, key)

    # Above this batch size the chunked copy-engine pipeline (soccer_step_host) matches or beats the zero-copy kernel
    # for the 3-streams-up / 3-streams-down formats (2^24 envs: 10.0-10.6 vs 9.9 G env-steps/s narrow, 5.5 vs 5.3 wide);
    # below it zero copy wins by up to 2x; the packed format is fastest zero-copy at every size
    # (profiles/r01g_time_host_paths.log)
    ZERO_COPY_MAX_ENVS = 1 << 23

    def step_host(self, act_a: Optional[torch.Tensor], act_b: Optional[torch.Tensor] = None,
                  rng8: Optional[torch.Tensor] = None, narrow: bool = False, n_chunks: int = 8, sync: bool = True,
                  zero_copy: Optional[bool] = None, rng32: Optional[torch.Tensor] = None,
                  rngf64: Optional[torch.Tensor] = None):
        """step() with HOST buffers -- the end-to-end path bench.py reports as `e2e`.

        In: uint8 CPU tensors (pinned memory keeps the copies asynchronous).  Out: CPU tensors
        owned by the env and overwritten by the next call -- obs int32 / reward float32 / flags
        uint8, or with narrow=True obs uint16 (viewed as int16 by torch) / reward int8: same values,
        4 instead of 9 bytes per env over PCIe.  The batch is cut into n_chunks slices whose upload,
        kernel and download overlap on three streams (soccer_step_host in the C ABI).  Batches of up to
        ZERO_COPY_MAX_ENVS envs with pinned inputs skip the staging: the kernel (wide or fused-narrow outputs) reads
        and writes the pinned host buffers directly (zero_copy=None: automatic).  alloc_host_inputs() hands out
        input buffers from huge-page-backed pinned memory.

        Every other mode of the env -- slip_prob > 0 (rng32 / rngf64 host tensors carry the step draw), the
        single-agent modes (the folded player's action tensor is None) and rng_mode='philox' (no draw tensors at all)
        -- runs zero copy with the natural dtypes: the step kernel of that mode reads the pinned input tensors and
        writes pinned int32 / float32 / uint8 result tensors itself, one launch and one synchronize per step."""
        if not self._plain:
            if narrow:
                raise NotImplementedError("narrow host streams exist for the plain step only (multi-agent, slip 0, injected)")
            common, hb = self._host_buffers(False)
            if self.num_envs == 0:
                return hb["h_obs"], hb["h_reward"], common["h_flags"]
            cur = torch.cuda.current_stream(self.device)
            self.step(act_a, act_b, rng8, rng32=rng32, rngf64=rngf64,
                      out=(hb["h_obs"], hb["h_reward"], common["h_flags"], None), _host=True)
            if sync:
                cur.synchronize()
            return hb["h_obs"], hb["h_reward"], common["h_flags"]
        for name, t in (("act_a", act_a), ("act_b", act_b), ("rng8", rng8)):
            if not (isinstance(t, torch.Tensor) and not t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()
                    and t.numel() == self.num_envs):
                raise ValueError(f"{name} must be a contiguous uint8 CPU tensor with {self.num_envs} elements")
        common, hb = self._host_buffers(narrow)
        if self.num_envs == 0:
            return hb["h_obs"], hb["h_reward"], common["h_flags"]
        if zero_copy is None:
            zero_copy = self.num_envs <= self.ZERO_COPY_MAX_ENVS
        if zero_copy and act_a.is_pinned() and act_b.is_pinned() and rng8.is_pinned():
            # the step kernel reads the pinned action / draw buffers and writes the pinned result buffers itself over
            # PCIe (unified addressing): one launch + one synchronize, no copies, no staging
            with torch.cuda.device(self.device):
                cur = torch.cuda.current_stream(self.device)
                st = C.c_void_p(cur.cuda_stream)
                tbl = None if self.table is None else _ptr(self.table)
                if narrow:
                    check(self.lib.soccer_step_narrow(C.byref(self.pitch), tbl, _ptr(self.state), _ptr(act_a), _ptr(act_b),
                                                      _ptr(rng8), _ptr(hb["h_obs"]), _ptr(hb["h_reward"]),
                                                      _ptr(common["h_flags"]), self.num_envs, st), "soccer_step_narrow")
                else:
                    ptrs = (_ptr(act_a), _ptr(act_b), _ptr(rng8), _ptr(hb["h_obs"]), _ptr(hb["h_reward"]),
                            _ptr(common["h_flags"]), None, self.num_envs, st)
                    if self.kernel == "table":
                        check(self.lib.soccer_step_table(C.byref(self.pitch), tbl, _ptr(self.state), *ptrs),
                              "soccer_step_table")
                    else:
                        check(self.lib.soccer_step(C.byref(self.pitch), _ptr(self.state), *ptrs), "soccer_step")
                self.step_count += 1
                if sync:
                    cur.synchronize()
            return hb["h_obs"], hb["h_reward"], common["h_flags"]
        s_in, s_k, s_out = common["streams"]
        cur = torch.cuda.current_stream(self.device)
        s_k.wait_stream(cur)                       # state may have been touched on the caller's stream
        a = _lib.StepHostArgs()
        a.state, a.table = self.state.data_ptr(), (None if self.table is None else self.table.data_ptr())
        a.scratch = common["scratch"].data_ptr()
        a.h_act_a, a.h_act_b, a.h_rng8 = act_a.data_ptr(), act_b.data_ptr(), rng8.data_ptr()
        a.h_obs, a.h_reward, a.h_flags = hb["h_obs"].data_ptr(), hb["h_reward"].data_ptr(), common["h_flags"].data_ptr()
        a.n, a.narrow, a.n_chunks = self.num_envs, int(bool(narrow)), max(1, int(n_chunks))
        a.s_in, a.s_compute, a.s_out = s_in.cuda_stream, s_k.cuda_stream, s_out.cuda_stream
        with torch.cuda.device(self.device):
            check(self.lib.soccer_step_host(C.byref(self.pitch), C.byref(a)), "soccer_step_host")
        cur.wait_stream(s_k)                       # later work on the caller's stream sees the new state
        self.step_count += 1
        if sync:
            s_out.synchronize()
        return hb["h_obs"], hb["h_reward"], common["h_flags"]

    # ------------------------------------------------------------------ packed streams (2 bytes in, 2 bytes out per env)
    @staticmethod
    def pack_joint(act_a: torch.Tensor, act_b: torch.Tensor) -> torch.Tensor:
        """The joint action (aa, ab) -- the key of the reference's P[s] (SIM:181-185) -- in one byte: aa | ab << 4."""
        return act_a | (act_b << 4)

    @staticmethod
    def unpack_result(w: torch.Tensor):
        """(obs, reward, terminated, truncated) from the 16-bit result words of step_packed / step_host_packed
        (int16 tensor): obs = w & 0xFFF, reward = w >> 14 (arithmetic), bit 12 terminated, bit 13 truncated."""
        w = w.to(torch.int32)
        return w & 0xFFF, (w >> 14).to(torch.float32), (w >> 12) & 1 != 0, (w >> 13) & 1 != 0

    def _check_packed(self):
        if self.kernel != "table" or self.slip_prob != 0.0 or not self.multiagent:
            raise NotImplementedError("the packed step needs kernel='table' (5x4 / 6x4 pitch), slip_prob == 0 and the "
                                      "multi-agent mode")

    def _packed_call(self, joint_ptr, rng8_ptr, res_ptr, stream):
        """soccer_step_table_packed, or soccer_step_table_packed_philox for rng_mode='philox' (no draw stream)."""
        if self.rng_mode == "philox":
            check(self.lib.soccer_step_table_packed_philox(C.byref(self.pitch), _ptr(self.table), _ptr(self.state), joint_ptr,
                                                           self.seed, self.step_count, self.env_id_base, res_ptr,
                                                           self.num_envs, stream), "soccer_step_table_packed_philox")
        else:
            check(self.lib.soccer_step_table_packed(C.byref(self.pitch), _ptr(self.table), _ptr(self.state), joint_ptr,
                                                    rng8_ptr, res_ptr, self.num_envs, stream), "soccer_step_table_packed")

    def step_packed(self, joint: torch.Tensor, rng8: Optional[torch.Tensor] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """step() on device tensors with packed streams (soccer_step_table_packed): joint = aa | ab << 4 (uint8),
        rng8 as in step() (None with rng_mode='philox'); returns the int16 result words (see unpack_result).  Same
        transition, reward, flags and fused reset as step(); 12 instead of 20 bytes of HBM traffic per env-step."""
        self._check_packed()
        if out is None:
            if getattr(self, "_result16", None) is None:
                self._result16 = torch.empty(self.num_envs, dtype=torch.int16, device=self.device)
            out = self._result16
        if self.num_envs == 0:
            return out
        philox = self.rng_mode == "philox"
        with torch.cuda.device(self.device):
            self._packed_call(_ptr(self._check_vec(joint, torch.uint8, "joint")),
                              None if philox else _ptr(self._check_vec(rng8, torch.uint8, "rng8")),
                              _ptr(self._check_vec(out, torch.int16, "out")), _stream(self.device))
        self.step_count += 1
        return out

    def step_host_packed(self, joint: torch.Tensor, rng8: Optional[torch.Tensor] = None, n_chunks: int = 8,
                         sync: bool = True, zero_copy: Optional[bool] = None) -> torch.Tensor:
        """step_host() with packed streams: 2 bytes up (joint action byte, draw byte) and 2 bytes down (one int16
        result word, see unpack_result) per env over PCIe instead of 3 + 4.  Returns a pinned CPU int16 tensor owned
        by the env and overwritten by the next call.  rng_mode='philox': no draw stream -- 1 byte up, 2 down.
        zero_copy: True = the kernel reads and writes the pinned host buffers itself; False = copy engines both ways,
        pipelined over n_chunks slices; "out" = hybrid: the copy engine uploads slice by slice while each slice's kernel
        writes its result words straight to the host (posted PCIe writes run at the link rate, the kernel's PCIe reads
        do not); None = True, the fastest measured at every batch size (2^24 envs: zero copy 18.8 G env-steps/s, hybrid and
        staged 16.3 G with 4 slices -- the chunked uploads only reach 32 GB/s; profiles/r02h_time_host_paths.log)."""
        self._check_packed()
        philox = self.rng_mode == "philox"
        if philox:
            rng8, zero_copy = None, True
        for name, t in (("joint", joint),) + (() if philox else (("rng8", rng8),)):
            if not (isinstance(t, torch.Tensor) and not t.is_cuda and t.dtype == torch.uint8 and t.is_contiguous()
                    and t.numel() == self.num_envs):
                raise ValueError(f"{name} must be a contiguous uint8 CPU tensor with {self.num_envs} elements")
        common, _ = self._host_buffers(True)
        if getattr(self, "_h_result16", None) is None:
            self._h_result16 = self._pinned(torch.int16)[0]
        h_res = self._h_result16
        if self.num_envs == 0:
            return h_res
        if zero_copy is None:
            zero_copy = True
        hybrid = zero_copy == "out"
        with torch.cuda.device(self.device):
            if zero_copy is True and joint.is_pinned() and (philox or rng8.is_pinned()):
                cur = torch.cuda.current_stream(self.device)
                self._packed_call(_ptr(joint), _ptr(rng8), _ptr(h_res), C.c_void_p(cur.cuda_stream))
                self.step_count += 1
                if sync:
                    cur.synchronize()
                return h_res
            if philox:
                raise ValueError("step_host_packed with rng_mode='philox' needs a pinned joint tensor (alloc_host_inputs)")
            s_in, s_k, s_out = common["streams"]
            cur = torch.cuda.current_stream(self.device)
            s_k.wait_stream(cur)
            a = _lib.StepHostArgs()
            a.state, a.table, a.scratch = self.state.data_ptr(), self.table.data_ptr(), common["scratch"].data_ptr()
            a.h_act_a, a.h_act_b, a.h_rng8 = joint.data_ptr(), None, rng8.data_ptr()
            a.h_obs, a.h_reward, a.h_flags = h_res.data_ptr(), None, None
            a.n, a.narrow, a.n_chunks = self.num_envs, HOST_PACKED, max(1, int(n_chunks))
            a.s_in, a.s_compute, a.s_out = s_in.cuda_stream, s_k.cuda_stream, s_out.cuda_stream
            a.d2h_zero_copy = 1 if hybrid else 0
            check(self.lib.soccer_step_host(C.byref(self.pitch), C.byref(a)), "soccer_step_host")
            cur.wait_stream(s_k)
            self.step_count += 1
            if sync:
                (s_k if hybrid else s_out).synchronize()
        return h_res

    def rollout(self, K: int, policy_a=None, policy_b=None, want_streams: bool = True, stats: Optional[torch.Tensor] = None,
                out=None):
        """K fused steps with on-device Philox draws and a uniform or table policy (K2).

        Returns (obs[K,N], reward[K,N], flags[K,N], stats[6]); stats accumulates
        [episodes, goals_A, goals_B, truncations, steps, sum_episode_len].  The reward stream is the env's return
        agent's, exactly what step() returns: player B's (negated) for an env constructed with player_a_policy
        (SIM:243-244); goals_A / goals_B always count by player A's sign.
        """
        n, dev = self.num_envs, self.device
        K = int(K)
        if out is not None:
            obs, reward, flags = out
            for name, t, dt in (("obs", obs, torch.int32), ("reward", reward, torch.float32), ("flags", flags, torch.uint8)):
                if t is not None and not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == dev and t.dtype == dt
                                          and t.is_contiguous() and t.numel() == K * n):
                    raise ValueError(f"out {name} must be a contiguous {dt} CUDA tensor of shape ({K}, {n})")
        elif want_streams:
            obs = torch.empty((K, n), dtype=torch.int32, device=dev)
            reward = torch.empty((K, n), dtype=torch.float32, device=dev)
            flags = torch.empty((K, n), dtype=torch.uint8, device=dev)
        else:
            obs = reward = flags = None
        if stats is None:
            stats = torch.zeros(6, dtype=torch.int64, device=dev)
        pa = self._policy_tensor(policy_a) if policy_a is not None else self.policy_a
        pb = self._policy_tensor(policy_b) if policy_b is not None else self.policy_b
        flip = 1 if self.policy_a is not None else 0          # the env's return agent is player_b (SIM:243-244)
        if n == 0 or K == 0:
            return obs, reward, flags, stats
        with torch.cuda.device(dev):
            st = _stream(dev)
            if self.kernel == "table":
                use_index = self.slip_index is not None and os.environ.get("SOCCER_B200_SLIP_WALK", "0") != "1"
                check(self.lib.soccer_rollout_table_policy(
                    C.byref(self.pitch), _ptr(self.table), _ptr(self.slip_index if use_index else None), _ptr(self.state),
                    _ptr(pa), _ptr(pb), self.seed, self.step_count, K, self.env_id_base, flip, _ptr(obs), _ptr(reward),
                    _ptr(flags), _ptr(stats), n, st), "soccer_rollout_table_policy")
            else:
                check(self.lib.soccer_rollout(C.byref(self.pitch), _ptr(self.state), _ptr(pa), _ptr(pb), self.seed,
                                              self.step_count, K, self.env_id_base, flip, _ptr(obs), _ptr(reward),
                                              _ptr(flags), _ptr(stats), n, st), "soccer_rollout")
        self.step_count += K
        return obs, reward, flags, stats
