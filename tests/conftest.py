import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: imports /root/reference (build container only)")


def golden_tags(kind):
    """['5x4_s000_multi', ...] for kind in {'table', 'rollout'}."""
    return sorted(os.path.basename(p)[len(f"ref_{kind}_"):-4]
                  for p in glob.glob(os.path.join(GOLDEN, f"ref_{kind}_*.npz")))


def load_golden(kind, tag):
    return np.load(os.path.join(GOLDEN, f"ref_{kind}_{tag}.npz"))


def parse_tag(tag):
    wh, s, mode = tag.split("_", 2)
    w, h = wh.split("x")
    return int(w), int(h), int(s[1:]) / 100.0, mode


@pytest.fixture(scope="session")
def oracle():
    from oracle import soccer_oracle
    soccer_oracle.build()
    return soccer_oracle
