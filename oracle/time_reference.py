"""Time the UNMODIFIED Python reference's step loop (BASELINE config 1) on this host.

MEASUREMENT INFRASTRUCTURE: imported by bench.py's cpu_baseline leg only.  Runs the staged copy of the reference
(oracle/_ref/reference_src.zip, see oracle/stage_ref.py; imported from the archive) in a subprocess pinned to one core, exactly as BASELINE.md's CPU-baseline plan
says: SoccerSimultaneousEnv(5, 4, slip_prob, seed=0) of the reference, actions from
np.random.RandomState(123).randint(0, 5, (steps, 2)), reset() whenever terminated or truncated, time.perf_counter.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
ZIP = os.path.join(REF, "reference_src.zip")

_CHILD = r'''
import json, sys, time
sys.path.insert(0, sys.argv[1])
import numpy as np
from gym_soccer.envs.soccer_simultaneous_env import SoccerSimultaneousEnv
steps, slip = int(sys.argv[2]), float(sys.argv[3])
t0 = time.perf_counter()
env = SoccerSimultaneousEnv(width=5, height=4, slip_prob=slip)
t_ctor = time.perf_counter() - t0
acts = np.random.RandomState(123).randint(0, 5, (steps, 2))
env.reset(seed=0)
ep = wins_a = wins_b = trunc = 0
t0 = time.perf_counter()
for aa, ab in acts:
    obs, rew, done, tr, info = env.step({'player_a': int(aa), 'player_b': int(ab)})
    if done['player_a'] or tr['player_a']:
        ep += 1
        wins_a += rew['player_a'] > 0
        wins_b += rew['player_a'] < 0
        trunc += (not done['player_a'])
        env.reset()
el = time.perf_counter() - t0
print(json.dumps({"steps": steps, "seconds": el, "steps_per_s": steps / el, "constructor_s": t_ctor, "episodes": int(ep),
                  "wins_a": int(wins_a), "wins_b": int(wins_b), "truncated": int(trunc), "slip_prob": slip}))
'''


def available() -> bool:
    return os.path.isfile(ZIP)


def time_python_reference(steps: int = 100000, slip_prob: float = 0.0, timeout_s: float = 120.0):
    """{'steps_per_s': ..., ...} of the staged reference, or None when oracle/_ref is absent."""
    if not available():
        return None
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    cmd = [sys.executable, "-c", _CHILD, ZIP, str(int(steps)), str(float(slip_prob))]
    try:
        out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout_s)
        rec = json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}
    try:
        with open(os.path.join(REF, "MANIFEST.json")) as fh:
            rec["manifest_files"] = len(json.load(fh)["sha256"])
    except OSError:
        pass
    rec["what"] = ("the unmodified reference (oracle/_ref, staged by oracle/stage_ref.py) behind the gym stand-in: "
                   "SoccerSimultaneousEnv(5, 4).step() loop, RandomState(123) actions, reset() on done / truncated, 1 core")
    return rec


if __name__ == "__main__":
    print(json.dumps(time_python_reference(int(sys.argv[1]) if len(sys.argv) > 1 else 100000), indent=1))
