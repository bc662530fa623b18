"""Import the UNMODIFIED reference (earasteh/Python-Motionplanning) for oracle pinning.

TEST INFRASTRUCTURE.  ``/root/reference`` exists only in the build container; on the GPU box the loader
falls back to the verbatim copy that ``baseline/stage_reference.py`` staged under ``baseline/_ref/``
(git-ignored, shipped with the snapshot).  Nothing at run time reads ``/root/reference`` on the box.

The reference does not import as-is in this image:
  * ``libs/vehicle_model/vehicle_model.py:7`` and ``libs/utils/env.py:8`` import
    ``matplotlib.pyplot`` (not installed)            -> stub modules in ``sys.modules``
  * ``libs/motionplanner/path_optimizer.py:172-173`` calls ``scipy.integrate.cumtrapz``
    (removed from scipy)                              -> alias to ``cumulative_trapezoid``
  * ``libs/vehicle_model/drive.py:153`` clears the terminal through ``os.system``
                                                      -> silenced on request
No reference file is edited or copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _has_reference(root: str) -> bool:
    return os.path.isfile(os.path.join(root, "libs", "vehicle_model", "vehicle_model.py"))


def _pick_root() -> str:
    env = os.environ.get("B200MP_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if _has_reference(cand):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return _has_reference(REFERENCE_ROOT)


def _install_shims() -> None:
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            pyplot = types.ModuleType("matplotlib.pyplot")
            anim = types.ModuleType("matplotlib.animation")
            anim.FuncAnimation = object
            mpl.pyplot = pyplot
            mpl.animation = anim
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = pyplot
            sys.modules["matplotlib.animation"] = anim
    import scipy.integrate

    if not hasattr(scipy.integrate, "cumtrapz"):
        scipy.integrate.cumtrapz = scipy.integrate.cumulative_trapezoid


class Reference(types.SimpleNamespace):
    """Handles to the reference modules on and around the hot path."""


def load(quiet_terminal: bool = True) -> Reference:
    if not available():
        raise RuntimeError(
            f"reference not found under {REFERENCE_ROOT} (nor staged under baseline/_ref: run baseline/stage_reference.py)")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ref = Reference()
    ref.vehicle_model = importlib.import_module("libs.vehicle_model.vehicle_model")
    ref.collision_checker = importlib.import_module("libs.motionplanner.collision_checker")
    ref.path_optimizer = importlib.import_module("libs.motionplanner.path_optimizer")
    ref.local_planner = importlib.import_module("libs.motionplanner.local_planner")
    ref.env = importlib.import_module("libs.utils.env")
    ref.drive = importlib.import_module("libs.vehicle_model.drive")
    ref.stanley = importlib.import_module("libs.controllers.stanley_controller")
    if quiet_terminal:
        ref.drive.os.system = lambda *a, **k: 0
    return ref
