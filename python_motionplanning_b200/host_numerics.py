"""Host-arithmetic probes that the bit-exact contract depends on.

``select_best_path_index`` in the reference scores paths with ``np.linalg.norm([a, b])``
(collision_checker.py:175, 188), i.e. ``sqrt(x.dot(x))`` through the host BLAS ``ddot``.  Whether that
dot product contracts into an FMA -- and in which order -- depends on the BLAS kernel chosen for the
host CPU (OpenBLAS Haswell/SkylakeX give ``sqrt(fma(v1, v1, v0*v0))``).  To return
the index the reference would return *on this host*, the engine asks the host once which closed form
its numpy follows and passes that to the kernel (``norm_mode`` of ``b200mp_select_best_f64``).

The candidates are evaluated in exact rational arithmetic (``fractions.Fraction`` -> one correct
rounding), so the probe needs no FMA instruction and no compiled helper.
"""
from __future__ import annotations

import math
import warnings
from fractions import Fraction
from functools import lru_cache

import numpy as np

from ._lib import NORM2_FMA_V0, NORM2_FMA_V1, NORM2_NOFMA


def norm2_closed_form(v0: float, v1: float, mode: int) -> float:
    if mode == NORM2_NOFMA:
        q = v0 * v0 + v1 * v1
    elif mode == NORM2_FMA_V1:
        q = float(Fraction(v1) * Fraction(v1) + Fraction(v0 * v0))
    elif mode == NORM2_FMA_V0:
        q = float(Fraction(v0) * Fraction(v0) + Fraction(v1 * v1))
    else:
        raise ValueError(f"unknown norm2 mode {mode}")
    return math.sqrt(q)


@lru_cache(maxsize=1)
def host_norm2_mode(n: int = 300, seed: int = 11) -> int:
    """The closed form ``np.linalg.norm([a, b])`` follows on this host (one of ``NORM2_*``)."""
    rng = np.random.default_rng(seed)
    v = rng.uniform(-60.0, 60.0, size=(n, 2))
    lit = [float(np.linalg.norm([a, b])) for a, b in v]
    for mode in (NORM2_FMA_V1, NORM2_NOFMA, NORM2_FMA_V0):
        if all(norm2_closed_form(float(a), float(b), mode) == l for (a, b), l in zip(v, lit)):
            return mode
    warnings.warn("np.linalg.norm([a, b]) on this host matches none of the known closed forms; "
                  "select_best_path_index may differ from the reference on near-ties", RuntimeWarning)
    return NORM2_NOFMA
