"""ctypes binding of libsoccer_b200.so (the C ABI in include/soccer_b200.h).

There is NO CPU fallback: if the shared library is missing it is built with nvcc, and if
that fails, or a kernel entry point returns a CUDA error, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_DEFAULT_SO = os.path.join(_HERE, "libsoccer_b200.so")
# SOCCER_B200_LIB points at an alternative build of the SAME sources (A/B kernel experiments)
SO_PATH = os.environ.get("SOCCER_B200_LIB", _DEFAULT_SO)

ABI_VERSION = 2

EXPORTS = (
    "soccer_abi_version", "soccer_pitch_info_host", "soccer_pack_state_host", "soccer_unpack_state_host",
    "soccer_state_to_obs_host", "soccer_obs_to_state_host", "soccer_reset", "soccer_reset_philox",
    "soccer_set_state", "soccer_get_obs", "soccer_step", "soccer_step_philox", "soccer_step_ex",
    "soccer_rollout", "soccer_sweep", "soccer_dense", "soccer_build_step_table", "soccer_step_table",
    "soccer_rollout_table", "soccer_step_table_bytes_host", "soccer_convert_state", "soccer_step_stats",
    "soccer_step_host", "soccer_step_host_scratch_bytes_host", "soccer_step_many", "soccer_bench_stream_mix",
    "soccer_bench_rollout_probe", "soccer_rollout_table_policy", "soccer_step_table_slip",
    "soccer_bellman_q", "soccer_plan", "soccer_plan_workspace_bytes_host", "soccer_step_table_philox", "soccer_step_table_packed", "soccer_step_narrow", "soccer_host_alloc", "soccer_host_free", "soccer_slip_index_bytes_host", "soccer_build_slip_index",
    "soccer_step_table_packed_philox", "soccer_dense_q", "soccer_policy_eval", "soccer_policy_eval_workspace_bytes_host",
    "soccer_cluster_table_bytes_host", "soccer_build_cluster_table", "soccer_rollout_table_cluster",
    "soccer_stats_allreduce_p2p_bytes_host", "soccer_stats_allreduce_p2p", "soccer_slip_danger_host",
    "soccer_step_speculate",
)


class SoccerB200Error(RuntimeError):
    pass


class Pitch(C.Structure):
    """struct soccer_pitch: the reference constructor's width/height/slip_prob (SIM:35)."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("slip_prob", C.c_double)]


class PitchInfo(C.Structure):
    _fields_ = [("padded_width", C.c_int32), ("height", C.c_int32), ("n_field_cells", C.c_int32),
                ("nS", C.c_int32), ("nA", C.c_int32), ("n_goal_rows", C.c_int32),
                ("goal_rows", C.c_int32 * 3), ("n_isd", C.c_int32), ("isd_obs", C.c_int32 * 4),
                ("isd_state", C.c_uint32 * 4), ("isd_tuple", (C.c_int32 * 5) * 4),
                ("slip_combo_prob", C.c_double * 9)]


class StepArgs(C.Structure):
    _fields_ = [("state", C.c_void_p), ("act_a", C.c_void_p), ("act_b", C.c_void_p), ("rng8", C.c_void_p),
                ("rng32", C.c_void_p), ("rngf64", C.c_void_p), ("policy_a", C.c_void_p),
                ("policy_b", C.c_void_p), ("obs", C.c_void_p), ("reward", C.c_void_p),
                ("flags", C.c_void_p), ("reset_obs", C.c_void_p), ("n", C.c_int64),
                ("auto_reset", C.c_int32), ("use_philox", C.c_int32), ("detail", C.c_int32),
                ("narrow", C.c_int32), ("seed", C.c_uint64), ("step", C.c_uint64),
                ("env_id_base", C.c_uint64), ("stats", C.c_void_p), ("table", C.c_void_p),
                ("slip_index", C.c_void_p)]


class StepHostArgs(C.Structure):
    _fields_ = [("state", C.c_void_p), ("table", C.c_void_p), ("scratch", C.c_void_p), ("h_act_a", C.c_void_p),
                ("h_act_b", C.c_void_p), ("h_rng8", C.c_void_p), ("h_obs", C.c_void_p), ("h_reward", C.c_void_p),
                ("h_flags", C.c_void_p), ("n", C.c_int64), ("narrow", C.c_int32), ("n_chunks", C.c_int32),
                ("s_in", C.c_void_p), ("s_compute", C.c_void_p), ("s_out", C.c_void_p), ("d2h_zero_copy", C.c_int32)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/ for sm_100a with nvcc (cross-compiles without a GPU)."""
    if SO_PATH != _DEFAULT_SO:
        if not os.path.exists(SO_PATH):
            raise SoccerB200Error(f"SOCCER_B200_LIB={SO_PATH} does not exist")
        return SO_PATH
    srcs = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "soccer_b200.h"))
    stale = (not os.path.exists(SO_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs if os.path.exists(s))
    if force or stale:
        cmd = ["make", "-C", _CSRC] + (["-B"] if force else [])
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose:
            print(r.stdout)
        if r.returncode != 0 or not os.path.exists(SO_PATH):
            raise SoccerB200Error("building libsoccer_b200.so failed (there is no CPU fallback):\n" + r.stdout)
    return SO_PATH


_lib = None


def lib():
    """The loaded library; builds it first if needed.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = build()
    try:
        L = C.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise SoccerB200Error(f"cannot load {path}: {e} (there is no CPU fallback)") from e
    missing = [n for n in EXPORTS if not hasattr(L, n)]
    if missing:
        raise SoccerB200Error(f"{path} lacks symbols {missing}")
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    PP = C.POINTER(Pitch)
    L.soccer_abi_version.restype = C.c_int
    L.soccer_abi_version.argtypes = []
    if L.soccer_abi_version() != ABI_VERSION:
        raise SoccerB200Error("libsoccer_b200.so ABI version mismatch")
    sig = {
        "soccer_pitch_info_host": [PP, C.POINTER(PitchInfo)],
        "soccer_pack_state_host": [PP, C.POINTER(C.c_int32 * 5), i32, i32, C.POINTER(C.c_uint32)],
        "soccer_unpack_state_host": [PP, C.c_uint32, C.POINTER(C.c_int32 * 5), C.POINTER(i32), C.POINTER(i32)],
        "soccer_state_to_obs_host": [PP, C.c_uint32, C.POINTER(i32)],
        "soccer_obs_to_state_host": [PP, i32, C.POINTER(C.c_uint32)],
        "soccer_reset": [PP, vp, vp, vp, vp, i64, vp],
        "soccer_reset_philox": [PP, vp, vp, vp, u64, u64, u64, i64, vp],
        "soccer_set_state": [PP, vp, vp, vp, i64, vp],
        "soccer_get_obs": [PP, vp, vp, i64, vp],
        "soccer_step": [PP, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp],
        "soccer_step_philox": [PP, vp, vp, vp, u64, u64, u64, vp, vp, vp, vp, i64, vp],
        "soccer_step_ex": [PP, C.POINTER(StepArgs), vp],
        "soccer_step_speculate": [PP, C.c_uint32, vp, vp, C.POINTER(C.c_double), vp, C.c_uint32, vp],
        "soccer_rollout": [PP, vp, vp, vp, u64, u64, i32, u64, i32, vp, vp, vp, vp, i64, vp],
        "soccer_sweep": [PP, i32, vp, vp, vp, vp, vp, vp],
        "soccer_dense": [PP, vp, vp, vp, vp, vp],
        "soccer_step_table_bytes_host": [PP, C.POINTER(i64)],
        "soccer_build_step_table": [PP, vp, vp],
        "soccer_step_table": [PP, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp],
        "soccer_step_table_philox": [PP, vp, vp, vp, vp, u64, u64, u64, vp, vp, vp, vp, i64, vp],
        "soccer_host_alloc": [C.c_size_t, C.POINTER(C.c_void_p)],
        "soccer_host_free": [C.c_void_p, C.c_size_t],
        "soccer_step_narrow": [PP, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp],
        "soccer_step_table_packed": [PP, vp, vp, vp, vp, vp, i64, vp],
        "soccer_rollout_table": [PP, vp, vp, u64, u64, i32, u64, vp, vp, vp, vp, i64, vp],
        "soccer_rollout_table_policy": [PP, vp, vp, vp, vp, vp, u64, u64, i32, u64, i32, vp, vp, vp, vp, i64, vp],
        "soccer_step_table_packed_philox": [PP, vp, vp, vp, u64, u64, u64, vp, i64, vp],
        "soccer_step_table_slip": [PP, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp],
        "soccer_slip_index_bytes_host": [PP, C.POINTER(i64)],
        "soccer_build_slip_index": [PP, vp, vp, vp],
        "soccer_convert_state": [PP, vp, vp, i32, i64, vp],
        "soccer_bellman_q": [PP, vp, vp, vp, C.c_double, vp, vp],
        "soccer_plan_workspace_bytes_host": [PP, C.POINTER(i64)],
        "soccer_plan": [PP, vp, vp, vp, C.c_double, C.c_double, i32, vp, vp, vp, vp, vp, vp],
        "soccer_step_stats": [vp, vp, i64, vp, vp],
        "soccer_slip_danger_host": [PP, C.POINTER(C.c_uint32 * 12), C.POINTER(i32)],
        "soccer_stats_allreduce_p2p_bytes_host": [C.POINTER(i64)],
        "soccer_stats_allreduce_p2p": [C.POINTER(C.c_uint64), i32, i32, u64, vp, vp],
        "soccer_cluster_table_bytes_host": [PP, C.POINTER(i64), C.POINTER(i32)],
        "soccer_build_cluster_table": [PP, vp, vp],
        "soccer_rollout_table_cluster": [PP, vp, vp, u64, u64, i32, u64, vp, vp, vp, vp, i64, vp],
        "soccer_dense_q": [PP, vp, vp, vp, C.c_double, vp, vp],
        "soccer_policy_eval_workspace_bytes_host": [PP, i32, C.POINTER(i64)],
        "soccer_policy_eval": [PP, vp, vp, vp, vp, C.c_double, C.c_double, i32, vp, vp, vp, vp],
        "soccer_step_host": [PP, C.POINTER(StepHostArgs)],
        "soccer_step_many": [PP, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, vp],
        "soccer_bench_stream_mix": [vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
        "soccer_bench_rollout_probe": [vp, i32, vp, vp, vp, i64, i32, vp],
        "soccer_step_host_scratch_bytes_host": [i64, C.POINTER(i64)],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    _lib = L
    return L


_ERR = {-1: "SOCCER_EINVAL (bad argument)", -2: "SOCCER_EPITCH (unsupported pitch)",
        -3: "SOCCER_ESLIP (slip_prob > 0 needs rng32/rngf64, or is unsupported by this entry point)",
        -4: "SOCCER_EPOLICY (both players cannot have a policy)",
        -5: "SOCCER_ETABLE (pitch too large for the shared-memory step table)"}


def check(rc: int, what: str):
    if rc == 0:
        return
    if rc < 0:
        raise SoccerB200Error(f"{what}: {_ERR.get(rc, rc)}")
    raise SoccerB200Error(f"{what}: CUDA error {rc} (no CPU fallback exists; a B200 is required)")


class HostArena:
    """One pinned, device-mapped host allocation backed by 2 MB huge pages (soccer_host_alloc), handed out as
    4 KB-aligned torch CPU tensors.  DMA reads from it run at the PCIe rate, which small individual
    `tensor.pin_memory()` allocations do not reliably reach (profiles/r01g_probe_pcie2.log).  The memory is
    released when the last tensor carved from it is gone."""

    def __init__(self, nbytes: int):
        import weakref
        import numpy as np
        import torch
        nbytes = max(int(nbytes), 1)
        ptr = C.c_void_p()
        L = lib()
        check(L.soccer_host_alloc(nbytes, C.byref(ptr)), "soccer_host_alloc")
        self.nbytes, self.used = nbytes, 0
        base = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(ptr.value))
        fin = weakref.finalize(base, L.soccer_host_free, ptr.value, nbytes)
        fin.atexit = False                      # at interpreter exit the CUDA context may already be gone
        self._bytes = torch.from_numpy(base)    # the storage keeps `base` alive

    def take(self, numel: int, dtype):
        """A pinned tensor of `numel` elements of `dtype` from the arena (4 KB aligned)."""
        import torch
        size = int(numel) * torch.empty(0, dtype=dtype).element_size()
        off = (self.used + 4095) // 4096 * 4096
        if off + size > self.nbytes:
            raise SoccerB200Error(f"HostArena exhausted: {off + size} > {self.nbytes} bytes")
        self.used = off + size
        return self._bytes[off:off + size].view(dtype)


def pitch_info(width: int, height: int, slip_prob: float = 0.0) -> PitchInfo:
    """Host-only: constructor products (SIM:48-65, 146-165).  Needs no GPU."""
    p = Pitch(int(width), int(height), float(slip_prob))
    out = PitchInfo()
    check(lib().soccer_pitch_info_host(C.byref(p), C.byref(out)), "soccer_pitch_info_host")
    return out
