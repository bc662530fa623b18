"""``install()``: plug the GPU engine into an unmodified checkout of the reference.

The reference looks its hot-path classes up through three module-level names, so rebinding them is
all the integration needs -- no reference file is edited:
  1. ``libs.vehicle_model.drive.VehicleModel``            (drive.py:3, used at :109)
  2. ``libs.motionplanner.local_planner.ThreadPool``      (local_planner.py:15, used at :370)
  3. ``libs.motionplanner.collision_checker.CollisionChecker`` (collision_checker.py:16, constructed by
     ``LocalPlanner.__init__`` through the ``collision_checker`` module)
Call it after importing the reference modules and before constructing ``Car`` / ``LocalPlanner``.
"""
from __future__ import annotations

import importlib
import sys

from .collision_checker import CollisionChecker
from .pool import ThreadPool
from .vehicle_model import VehicleModel, VehicleParameters

_SEAMS = (
    ("libs.vehicle_model.drive", "VehicleModel", VehicleModel),
    ("libs.motionplanner.local_planner", "ThreadPool", ThreadPool),
    ("libs.motionplanner.collision_checker", "CollisionChecker", CollisionChecker),
)
_saved = {}


def install(import_missing: bool = True):
    """Rebind the three seams; returns the list of ``module.name`` strings that were rebound."""
    done = []
    for mod_name, attr, repl in _SEAMS:
        mod = sys.modules.get(mod_name)
        if mod is None and import_missing:
            try:
                mod = importlib.import_module(mod_name)
            except ImportError:
                mod = None
        if mod is None:
            continue
        key = (mod_name, attr)
        if key not in _saved:
            _saved[key] = getattr(mod, attr)
        setattr(mod, attr, repl)
        done.append(f"{mod_name}.{attr}")
    return done


def uninstall():
    """Restore the reference's own classes."""
    for (mod_name, attr), orig in list(_saved.items()):
        mod = sys.modules.get(mod_name)
        if mod is not None:
            setattr(mod, attr, orig)
        del _saved[(mod_name, attr)]


__all__ = ["install", "uninstall", "VehicleModel", "VehicleParameters", "CollisionChecker", "ThreadPool"]
