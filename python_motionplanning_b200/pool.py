"""The fan-out seam: a stand-in for the reference's ``multiprocessing.Pool``.

The reference forks one process per candidate path on every planner call
(``local_planner.py:15`` imports ``Pool`` as ``ThreadPool``; ``:369-372`` runs
``ThreadPool(processes=len(paths)).starmap(self._collision_checker.collision_check, zip(paths, repeat(obstacle)))``;
``:373-374`` turns the ``ValueError`` for zero paths into ``[True] * 7``).  This class keeps that
constructor / ``starmap`` contract, recognises a bound ``collision_check`` and issues ONE batched kernel
launch for all paths instead.
"""
from __future__ import annotations


class ThreadPool:
    def __init__(self, processes=None, *args, **kwargs):
        if processes is not None and processes < 1:
            # multiprocessing.Pool raises exactly this; the planner's `except ValueError` depends on it
            raise ValueError("Number of processes must be at least 1")
        self._processes = processes

    def starmap(self, func, iterable, chunksize=None):
        tasks = [tuple(a) for a in iterable]
        owner = getattr(func, "__self__", None)
        batch = getattr(owner, "collision_check_paths", None)
        if batch is not None and getattr(func, "__name__", "") == "collision_check" and tasks:
            obstacles = tasks[0][1]
            if all(len(t) == 2 and t[1] is obstacles for t in tasks):
                return batch([t[0] for t in tasks], obstacles)
            return [batch([t[0]], t[1])[0] for t in tasks]
        # anything else is host-side plumbing, run in-process in order
        return [func(*t) for t in tasks]

    def map(self, func, iterable, chunksize=None):
        return self.starmap(func, ((a,) for a in iterable))

    def close(self):
        pass

    def join(self):
        pass

    def terminate(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
