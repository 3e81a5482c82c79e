#!/usr/bin/env python
"""Stages the UNMODIFIED reference tree under the git-ignored ``baseline/_ref/`` so that it travels to the GPU box.

The reference (17 loose files, no setup.py / pyproject.toml) cannot be ``pip install --target``-ed; this script is the
equivalent install step: a byte-for-byte copy of the reference's Python files, nothing edited, nothing added except a
MANIFEST with their sha256. ``baseline/_ref/`` is listed in .gitignore (the sources never enter this repo's history)
but not in .gpurunignore. It is used ONLY as the other side of comparisons:

  * tests/test_trainer_gpu.py runs the reference's own ``utils/trainer.py`` (train_one_epoch / validate / test) with
    ``models.*`` resolving to THIS repo's drop-in modules — the "drops into utils/trainer.py unchanged" claim;
  * ``bench.py --impl reference`` / ``cpu_baseline`` time the reference's own ``models/model.py`` + ``models/loss.py``
    on the host cores (kind "reference"); ``bench.py``'s ``gpu_baseline`` leg runs the same modules on stock PyTorch.

Run in the build container (``/root/reference`` exists only there): ``python tools/stage_reference.py``.
``__graft_entry__.build()`` calls it when the reference tree is present.
"""
import hashlib
import json
import os
import shutil
import sys

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["models/model.py", "models/loss.py", "models/vnet.py", "models/mod.py", "utils/trainer.py", "utils/utils.py",
         "test.py"]


def stage(ref=REF, dst=DST):
    if not os.path.isdir(ref):
        return False
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref, rel)
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(src, out)
        with open(out, "rb") as f:
            manifest[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref, "note": "unmodified copies; see tools/stage_reference.py", "sha256": manifest}, f,
                  indent=1)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else "reference tree not present", DST)
    sys.exit(0 if ok else 1)
