"""Stage the UNMODIFIED Python reference so that it can be TIMED on the GPU box's host cores.

    python oracle/stage_ref.py          # build container only: copies into oracle/_ref/ (git-ignored, ships with gpurun)

TEST / MEASUREMENT INFRASTRUCTURE.  /root/reference does not exist on the GPU box, and the reference is pure Python
(no native build), so `__graft_entry__.build()` -- which runs where /root/reference exists -- copies the files of the
reference's step path, byte for byte, next to the 4-module `gym` stand-in (oracle/ref_shim: the three `gym` names the
reference imports; `gym` itself is not in the image), into oracle/_ref/:

    oracle/_ref/reference_src.zip   gym_soccer/... <- /root/reference/gym_soccer (package sources only, tests excluded)
                                    gym/...        <- oracle/ref_shim/gym
                                    (imported straight from the archive: zipimport)
    oracle/_ref/MANIFEST.json       sha256 of every staged reference file

oracle/_ref/ is listed in .gitignore: no reference source enters the repository's history.  Only bench.py's
cpu_baseline leg imports it (oracle/time_reference.py); nothing under gym_soccer_littman94_b200/ does.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SOCCER_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def stage() -> bool:
    src = os.path.join(REFERENCE_ROOT, "gym_soccer")
    if not os.path.isfile(os.path.join(src, "envs", "soccer_simultaneous_env.py")):
        return False
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    manifest = {}
    with zipfile.ZipFile(os.path.join(DEST, "reference_src.zip"), "w", zipfile.ZIP_DEFLATED) as z:
        for dirpath, dirnames, files in os.walk(src):
            dirnames[:] = sorted(d for d in dirnames if d not in ("tests", "__pycache__"))
            for f in sorted(files):
                if not f.endswith(".py"):
                    continue
                full = os.path.join(dirpath, f)
                rel = os.path.relpath(full, REFERENCE_ROOT)
                z.write(full, rel)
                manifest[rel] = hashlib.sha256(open(full, "rb").read()).hexdigest()
        shim = os.path.join(HERE, "ref_shim")
        for dirpath, dirnames, files in os.walk(os.path.join(shim, "gym")):
            dirnames[:] = sorted(d for d in dirnames if d != "__pycache__")
            for f in sorted(files):
                if f.endswith(".py"):
                    z.write(os.path.join(dirpath, f), os.path.relpath(os.path.join(dirpath, f), shim))
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as fh:
        json.dump({"reference_root": REFERENCE_ROOT, "sha256": manifest}, fh, indent=1)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged the reference into", DEST if ok else "(nothing: reference not found)")
    sys.exit(0 if ok else 1)
