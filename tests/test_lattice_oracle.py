"""CPU: the spiral sampling / frame transform restatement (oracle/lattice_numpy.py) and the
product-side vectorised generator (workloads.sample_spirals / transform_to_global) against the literal reference
(tests/golden/lattice_paths.npz, made by oracle/make_golden.py --only lattice)."""
import numpy as np

from oracle import lattice_numpy as ln
from python_motionplanning_b200 import workloads as wl


def test_sample_spiral_and_transform_vs_literal(golden):
    g = golden("lattice_paths.npz")
    for i in range(len(g["sf"])):
        sp = ln.sample_spiral([g["kappa1"][i], g["kappa2"][i], g["sf"][i]])
        assert (len(sp[0]), len(sp[1]), len(sp[2])) == (49, 49, 50)
        assert np.array_equal(sp[0], g["x"][i]) and np.array_equal(sp[1], g["y"][i]) and np.array_equal(sp[2], g["t"][i])
        tp = ln.transform_paths([sp], list(g["ego"][i]))[0]
        assert np.array_equal(tp[0], g["gx"][i]) and np.array_equal(tp[1], g["gy"][i]) and np.array_equal(tp[2], g["gt"][i])


def test_vectorised_generator_vs_literal(golden):
    g = golden("lattice_paths.npz")
    x, y, t = wl.sample_spirals(g["kappa1"], g["kappa2"], g["sf"])
    assert np.abs(x - g["x"]).max() < 1e-12 and np.abs(y - g["y"]).max() < 1e-12 and np.abs(t - g["t"]).max() < 1e-13
    gx, gy, gt = wl.transform_to_global(x, y, t, g["ego"][:, 0], g["ego"][:, 1], g["ego"][:, 2])
    assert np.abs(gx - g["gx"]).max() < 1e-11 and np.abs(gy - g["gy"]).max() < 1e-11 and np.abs(gt - g["gt"]).max() < 1e-13
