"""The statistics vector bench.py's `shard_invariant` check expects, computed by the CPU ORACLE (test infrastructure).

    python oracle/make_shard_invariant.py          # ~1 minute on 8 cores; prints the constant recorded in bench.py

Workload: 2^24 GLOBAL envs (5x4, slip 0), Philox seed 20261018, reset() then K = 64 lock-steps of uniform play.  Philox is
keyed by the global env id, so the statistics are a pure function of this description -- whatever the number of GPUs
the batch is sharded over.  bench.py runs it sharded N ways at every N and asserts that the all-reduced vector equals
this one.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import soccer_oracle as so  # noqa: E402

N, K, SEED = 1 << 24, 64, 20261018


def main():
    m = so.OracleModel(5, 4, 0.0)
    L = so.lib()
    # reset draw of env i = bits 2..3 of word(seed, i, 2^64 - 1)
    table = np.zeros(4, so.STATE_DTYPE)
    for k in range(4):
        table[k] = m.isd[k][1]
    sel = np.empty(N, np.int64)
    step = (1 << 64) - 1
    for i in range(N):
        sel[i] = (L.orc_philox_word(SEED, i, step) >> 2) & 3
    states = table[sel].copy()
    ts = np.zeros(N, np.int32)
    _, _, _, st = m.rollout_philox(states, ts, K, SEED, n_threads=os.cpu_count() or 1, want_streams=False)
    print("SHARD_INVARIANT =", dict(envs=N, K=K, seed=SEED, stats=[int(x) for x in st]))


if __name__ == "__main__":
    main()
