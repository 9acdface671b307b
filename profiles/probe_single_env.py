"""Single-env drop-in: step() calls per second, speculative step vs launch-and-wait, slip 0 and 0.2 (BASELINE config 1's loop)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_soccer_littman94_b200.envs import SoccerSimultaneousEnv
dev = torch.device("cuda", 0)
for slip in (0.0, 0.2):
    for mode in ("speculative", "zero_copy"):
        os.environ["SOCCER_B200_SINGLE_ENV_SPECULATE"] = "1" if mode == "speculative" else "0"
        e1 = SoccerSimultaneousEnv(5, 4, slip_prob=slip, seed=0, device=dev)
        acts = np.random.RandomState(123).randint(0, 5, (50000, 2))
        e1.reset()
        for aa, ab in acts[:2000]:
            _, _, dn, tr, _ = e1.step({'player_a': int(aa), 'player_b': int(ab)})
            if dn['player_a'] or tr['player_a']:
                e1.reset()
        t0 = time.perf_counter(); n_ep = 0
        for aa, ab in acts:
            _, _, dn, tr, _ = e1.step({'player_a': int(aa), 'player_b': int(ab)})
            if dn['player_a'] or tr['player_a']:
                e1.reset(); n_ep += 1
        dt = time.perf_counter() - t0
        print("slip %.1f %-12s %.0f steps/s  (%d episodes)" % (slip, mode, len(acts) / dt, n_ep), flush=True)
