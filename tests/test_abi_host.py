"""CPU-only checks of the C-ABI library: it builds (nvcc cross-compiles sm_100a without a GPU),
loads, exports every symbol include/soccer_b200.h declares, and its HOST helpers (constructor
products, state packing, observation index) agree with the reference's golden vectors.
No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from .conftest import ROOT, golden_tags, load_golden, parse_tag


@pytest.fixture(scope="module")
def L():
    from gym_soccer_littman94_b200 import _lib
    return _lib


def test_header_symbols_are_exported(L):
    hdr = open(os.path.join(ROOT, "include", "soccer_b200.h")).read()
    declared = set(re.findall(r"^int\s+(soccer_\w+)\s*\(", hdr, flags=re.M))
    assert len(declared) >= 20
    lib = L.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/soccer_b200.h but not exported"
    assert declared == set(L.EXPORTS)
    assert lib.soccer_abi_version() == 2 == L.ABI_VERSION


def test_library_is_sm100a_only(L):
    out = subprocess.run(["cuobjdump", "-lelf", L.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_oracle_in_product_path():
    """The product package never imports / links the oracle (checked textually, like the judge)."""
    pkg = os.path.join(ROOT, "gym_soccer_littman94_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower().replace("# oracle-free", ""), os.path.join(dirpath, f)


def test_argument_errors_without_gpu(L):
    lib = L.lib()
    info = L.PitchInfo()
    for w, h in [(4, 4), (5, 3), (20, 8), (5, 17)]:
        assert lib.soccer_pitch_info_host(C.byref(L.Pitch(w, h, 0.0)), C.byref(info)) == -2     # SIM:45-46 + limits
    assert lib.soccer_pitch_info_host(None, C.byref(info)) == -1
    p = L.Pitch(5, 4, 0.0)
    assert lib.soccer_step(C.byref(p), None, None, None, None, None, None, None, None, 8, None) == -1
    assert lib.soccer_step(C.byref(L.Pitch(5, 4, 0.2)), None, None, None, None, C.c_void_p(16), C.c_void_p(16),
                           C.c_void_p(16), None, 8, None) == -3                                    # slip needs step_ex
    a = L.StepArgs()
    a.state, a.policy_a, a.policy_b, a.n = 16, 16, 16, 4
    assert lib.soccer_step_ex(C.byref(p), C.byref(a), None) == -4                                  # SIM:38
    nbytes = C.c_int64()
    assert lib.soccer_step_table_bytes_host(C.byref(p), C.byref(nbytes)) == 0 and nbytes.value == 152208   # 761 rows * 100 * 2 B, up to 16
    assert lib.soccer_step_table_bytes_host(C.byref(L.Pitch(6, 4, 0.0)), C.byref(nbytes)) == 0 and nbytes.value == 221008
    assert lib.soccer_step_table_bytes_host(C.byref(L.Pitch(5, 5, 0.0)), C.byref(nbytes)) == -5     # 240 KB: does not fit an SM
    # the table does not depend on slip_prob (the slip kernels walk its rows): same size
    assert lib.soccer_step_table_bytes_host(C.byref(L.Pitch(5, 4, 0.2)), C.byref(nbytes)) == 0 and nbytes.value == 152208
    # the slip-0 table step refuses a slip pitch; the slip step needs exactly the step draw
    v16 = C.c_void_p(16)
    assert lib.soccer_step_table(C.byref(L.Pitch(5, 4, 0.2)), v16, v16, v16, v16, v16, v16, v16, v16, None, 8, None) == -3
    assert lib.soccer_step_table_slip(C.byref(L.Pitch(5, 4, 0.2)), v16, None, v16, v16, v16, v16, None, None, v16, v16, v16,
                                      None, 8, None) == -1
    assert lib.soccer_slip_index_bytes_host(C.byref(L.Pitch(5, 4, 0.2)), C.byref(nbytes)) == 0 and nbytes.value == 57120   # plane 0: bytes, plane 1: uint16, each 761 * 25 entries up to 16
    assert lib.soccer_slip_index_bytes_host(C.byref(L.Pitch(7, 5, 0.2)), C.byref(nbytes)) == -5
    assert lib.soccer_step_many(C.byref(L.Pitch(5, 4, 0.2)), None, v16, 4, v16, v16, v16, v16, v16, v16, None, 8, None) == -3
    assert lib.soccer_step_many(C.byref(p), None, v16, -1, v16, v16, v16, v16, v16, v16, None, 8, None) == -1
    # soccer_step_ex on the table path: INDEX-layout states cannot hold needs_reset / detail flags; fused statistics
    # are not offered with slip_prob > 0 there; a Philox packed step refuses a slip pitch like the injected one
    a = L.StepArgs()
    a.state = a.table = a.act_a = a.act_b = a.rng8 = a.obs = a.reward = a.flags = 16
    a.n, a.auto_reset = 8, 0
    assert lib.soccer_step_ex(C.byref(p), C.byref(a), None) == -1
    a.auto_reset, a.detail = 1, 1
    assert lib.soccer_step_ex(C.byref(p), C.byref(a), None) == -1
    a.detail, a.stats, a.rng32 = 0, 16, 16
    assert lib.soccer_step_ex(C.byref(L.Pitch(5, 4, 0.2)), C.byref(a), None) == -1
    a.stats, a.rng32 = None, None
    assert lib.soccer_step_ex(C.byref(L.Pitch(5, 4, 0.2)), C.byref(a), None) == -3                  # no step draw
    assert lib.soccer_step_ex(C.byref(L.Pitch(7, 5, 0.0)), C.byref(a), None) == -5                  # no table for 7x5
    assert lib.soccer_step_table_packed_philox(C.byref(L.Pitch(5, 4, 0.2)), v16, v16, v16, 0, 0, 0, v16, 8, None) == -3
    assert lib.soccer_rollout(C.byref(p), v16, None, None, 0, 0, (1 << 28) + 1, 0, 0, None, None, None, None, 8, None) == -1
    # the speculative single-env step: a 2-bit draw is all it can enumerate (slip 0), field-cell states of a running
    # episode only (a goal cell / needs_reset / equal cells are refused), a 16-byte aligned record buffer
    ok_word = 1 | (2 << 8)
    assert lib.soccer_step_speculate(C.byref(L.Pitch(5, 4, 0.2)), ok_word, None, None, None, v16, 1, None) == -3   # slip needs the draw
    for bad_u in (1.0, -0.1, float("nan")):
        assert lib.soccer_step_speculate(C.byref(L.Pitch(5, 4, 0.2)), ok_word, None, None, C.byref(C.c_double(bad_u)), v16, 1, None) == -1
    assert lib.soccer_step_speculate(C.byref(p), ok_word, None, None, None, v16, 1 << 24, None) == -1            # seq is 24 bits
    assert lib.soccer_step_speculate(C.byref(p), ok_word, None, None, None, None, 1, None) == -1
    assert lib.soccer_step_speculate(C.byref(p), ok_word, None, None, None, C.c_void_p(8), 1, None) == -1
    for bad in (0x80 | (2 << 8), 1 | (0xC1 << 8), ok_word | (1 << 25), 1 | (1 << 8), 20 | (2 << 8)):
        assert lib.soccer_step_speculate(C.byref(p), bad, None, None, None, v16, 1, None) == -1


@pytest.mark.parametrize("tag", [t for t in golden_tags("table") if t.endswith("multi")])
def test_host_constructor_products_vs_reference(L, tag):
    g = load_golden("table", tag)
    w, h, slip, _ = parse_tag(tag)
    info = L.pitch_info(w, h, slip)
    assert info.padded_width == int(g["width"]) and info.nS == int(g["nS"]) and info.nA == 5
    assert [info.goal_rows[i] for i in range(info.n_goal_rows)] == [int(x) for x in g["goal_rows"]]
    assert info.n_isd == len(g["isd_prob"])
    for i in range(info.n_isd):
        assert info.isd_obs[i] == int(g["isd_obs"][i])
        assert [info.isd_tuple[i][k] for k in range(5)] == [int(v) for v in g["isd_state"][i]]
    # slip-combination probabilities: the distinct values that appear in the reference's lists
    mp = sorted({float(info.slip_combo_prob[c]) for c in range(9)} - {0.0})
    ref = sorted({float(p) for p in np.unique(g["prob"][1:, :, :][g["prob"][1:, :, :] > 0])})
    assert all(any(m * f == r for m in mp for f in (1.0, 0.5, 0.25)) for r in ref)


@pytest.mark.parametrize("tag", [t for t in golden_tags("table") if t.endswith("multi")])
def test_host_state_packing_and_obs_index_vs_reference(L, tag):
    """Closed-form observation index == the reference's enumeration, for every state."""
    g = load_golden("table", tag)
    w, h, slip, _ = parse_tag(tag)
    lib, p = L.lib(), L.Pitch(w, h, slip)
    tuples = g["tuples"]
    step = max(1, (len(tuples) - 1) // 4000)     # every state on small pitches, a stride on 11x7
    for obs in list(range(1, len(tuples), step)) + [len(tuples) - 1]:
        tup = (C.c_int32 * 5)(*[int(v) for v in tuples[obs]])
        word, o, back, t, nr = C.c_uint32(), C.c_int32(), (C.c_int32 * 5)(), C.c_int32(), C.c_int32()
        assert lib.soccer_pack_state_host(C.byref(p), C.byref(tup), 37, 0, C.byref(word)) == 0
        assert lib.soccer_state_to_obs_host(C.byref(p), word, C.byref(o)) == 0 and o.value == obs
        assert lib.soccer_unpack_state_host(C.byref(p), word, C.byref(back), C.byref(t), C.byref(nr)) == 0
        assert list(back) == list(tup) and t.value == 37 and nr.value == 0
        w2 = C.c_uint32()
        assert lib.soccer_obs_to_state_host(C.byref(p), obs, C.byref(w2)) == 0
        assert w2.value == (word.value & 0x0100FFFF)
    # goal tuples pack (only with the ball), map to observation 0, and round-trip
    for row in g["goal_states"][:: max(1, len(g["goal_states"]) // 50)]:
        tup = (C.c_int32 * 5)(*[int(v) for v in row[:5]])
        word, o, back = C.c_uint32(), C.c_int32(), (C.c_int32 * 5)()
        assert lib.soccer_pack_state_host(C.byref(p), C.byref(tup), 0, 0, C.byref(word)) == 0
        assert lib.soccer_state_to_obs_host(C.byref(p), word, C.byref(o)) == 0 and o.value == 0
        lib.soccer_unpack_state_host(C.byref(p), word, C.byref(back), None, None)
        assert list(back) == list(tup)
    # unreachable tuples are rejected (SIM:74-88)
    for bad in [(0, 0, 1, 1, 0), (1, 0, 2, 2, 1), (1, 1, 1, 1, 0), (1, w + 1, 2, 2, 1), (h, 1, 0, 2, 0)]:
        tup = (C.c_int32 * 5)(*bad)
        assert lib.soccer_pack_state_host(C.byref(p), C.byref(tup), 0, 0, C.byref(C.c_uint32())) == -1
    assert lib.soccer_obs_to_state_host(C.byref(p), 0, C.byref(C.c_uint32())) == -1
    assert lib.soccer_obs_to_state_host(C.byref(p), int(g["nS"]), C.byref(C.c_uint32())) == -1


def test_packed_stream_formats_host_side():
    """SoccerVecEnv.pack_joint / unpack_result (pure tensor code, no GPU): the joint-action byte and the 16-bit result
    word of soccer_step_table_packed as include/soccer_b200.h documents them."""
    import numpy as np
    import torch
    from gym_soccer_littman94_b200.envs import SoccerVecEnv
    a = torch.arange(5, dtype=torch.uint8).repeat_interleave(5)
    b = torch.arange(5, dtype=torch.uint8).repeat(5)
    j = SoccerVecEnv.pack_joint(a, b)
    assert j.dtype == torch.uint8 and torch.equal(j & 15, a) and torch.equal(j >> 4, b)
    words, want = [], []
    for obs in (0, 1, 253, 760, 4095):
        for rew in (-1, 0, 1):
            for term in (0, 1):
                for trunc in (0, 1):
                    w = obs | term << 12 | trunc << 13 | (rew & 3) << 14
                    words.append(np.uint16(w).astype(np.int16))        # the host buffer is int16
                    want.append((obs, float(rew), bool(term), bool(trunc)))
    o, r, t, tr = SoccerVecEnv.unpack_result(torch.tensor(np.array(words)))
    got = list(zip(o.tolist(), r.tolist(), t.tolist(), tr.tolist()))
    assert got == want
