"""Stage the UNMODIFIED reference checkout under ``baseline/_ref/`` so it travels to the GPU box.

    python baseline/stage_reference.py [--source /root/reference] [--force]

``/root/reference`` exists only in the build container; ``baseline/_ref/`` is git-ignored (the reference's
sources never enter this repository's history) but NOT gpurun-ignored, so the snapshot that goes to the
B200 box carries a verbatim copy.  It is used there for two things only:
  * the config-1 test: the reference's own ``Car.drive`` loop (``libs/vehicle_model/drive.py:112-154``, as
    ``animate.py:27, 61-99`` drives it) run unchanged on the engine through ``install()``;
  * the literal CPU baseline: the reference's ``VehicleModel.planar_model_RK4``
    (``libs/vehicle_model/vehicle_model.py:427-445``) under ``multiprocessing.Pool(os.cpu_count())`` and its
    ``CollisionChecker.collision_check`` (``libs/motionplanner/collision_checker.py:32-117``), timed beside the GPU.
The copy is byte-for-byte (checked file by file below, and recorded in ``baseline/_ref/.staged.json``); only the
README's animation (``resources/``) is left out.  No product module imports anything from it.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SKIP_DIRS = {".git", "resources", "__pycache__"}


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _files(root: str):
    for d, dirs, files in os.walk(root):
        dirs[:] = sorted(x for x in dirs if x not in SKIP_DIRS)
        for f in sorted(files):
            if f.endswith(".pyc") or f == ".staged.json":
                continue
            p = os.path.join(d, f)
            yield os.path.relpath(p, root), p


def stage(source: str = "/root/reference", force: bool = False) -> str | None:
    """Copy ``source`` to ``baseline/_ref``; returns the destination, or None when there is no source
    (the GPU box: whatever was staged before stays as it is)."""
    if not os.path.isfile(os.path.join(source, "libs", "vehicle_model", "vehicle_model.py")):
        return None
    want = {rel: _sha(p) for rel, p in _files(source)}
    manifest = os.path.join(DEST, ".staged.json")
    if not force and os.path.isfile(manifest):
        try:
            have = json.load(open(manifest))["sha256"]
            if have == want and all(os.path.isfile(os.path.join(DEST, r)) and _sha(os.path.join(DEST, r)) == s
                                    for r, s in want.items()):
                return DEST
        except Exception:
            pass
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    for rel, p in _files(source):
        out = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(p, out)
        if _sha(out) != want[rel]:
            raise RuntimeError(f"copy of {rel} differs from the source")
    with open(manifest, "w") as f:
        json.dump({"source": source, "files": len(want), "sha256": want}, f, indent=1, sort_keys=True)
    return DEST


def verify() -> bool:
    """True when ``baseline/_ref`` holds exactly the files its manifest lists, unmodified."""
    manifest = os.path.join(DEST, ".staged.json")
    if not os.path.isfile(manifest):
        return False
    want = json.load(open(manifest))["sha256"]
    have = {rel: _sha(p) for rel, p in _files(DEST)}
    return have == want


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    d = stage(a.source, a.force)
    print(d if d else f"no reference under {a.source}; nothing staged")
    sys.exit(0 if (d or os.path.isdir(DEST)) else 1)
