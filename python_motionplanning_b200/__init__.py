"""python_motionplanning_b200 -- B200-native batched vehicle dynamics and lattice evaluation.

A drop-in for the one data-parallel hot path of earasteh/Python-Motionplanning (the 7-DoF planar
vehicle model step and the circle-offset collision test + best-path selection), built from scratch as
hand-written CUDA for sm_100a behind a C ABI (``include/b200mp.h``, ``libb200mp.so``).

  reference-named surface : VehicleParameters, VehicleModel, CollisionChecker, ThreadPool, install()
  batch surface           : Engine (rollout, rollout_to_host, planar_model_batch, collision_check_batch,
                            select_best_path_index_batch, mpc_sample_controls, argmin, fma_peak,
                            track_closed_loop -> DataLog rows, datalog.write_results_csv)
  multi-GPU               : distributed (shard rollouts / paths across ranks, gather costs, broadcast winner)

There is no CPU fallback: without the CUDA library and a GPU every compute call raises.
"""
from ._lib import B200mpError, load as load_library
from .collision_checker import CollisionChecker
from .engine import Engine, RolloutResult, TrackGains, TrackResult, pack_params
from .install import install, uninstall
from .pool import ThreadPool
from .vehicle_model import VehicleModel, VehicleParameters, default_engine

__all__ = ["B200mpError", "load_library", "CollisionChecker", "Engine", "RolloutResult", "TrackGains", "TrackResult", "pack_params", "install",
           "uninstall", "ThreadPool", "VehicleModel", "VehicleParameters", "default_engine"]
__version__ = "0.1.0"
