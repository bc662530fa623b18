"""CPU: the spiral-optimisation restatement (oracle/spiral_opt_numpy.py) against the literal
reference: objective and gradient values of ``PathOptimizer.objective`` / ``objective_grad`` at 256 random points, and
the minimisers scipy's L-BFGS-B returns for 192 goal states (tests/golden/spiral_opt.npz)."""
import numpy as np

from oracle import spiral_opt_numpy as so


def test_objective_and_gradient_vs_literal(golden):
    g = golden("spiral_opt.npz")
    f = so.objective(g["eval_p"], g["eval_goal"])
    gr = so.objective_grad(g["eval_p"], g["eval_goal"])
    assert (np.abs(f - g["eval_f"]) / np.maximum(np.abs(g["eval_f"]), 1.0)).max() < 1e-12
    assert (np.abs(gr - g["eval_grad"]) / np.maximum(np.abs(g["eval_grad"]), 1.0)).max() < 1e-11


def test_minimisers_vs_scipy_lbfgsb(golden):
    """Same basin, at least as deep: the restated solver's objective is never above the reference's result, and the
    parameters agree to the accuracy L-BFGS-B stops at."""
    g = golden("spiral_opt.npz")
    worst_p, worst_f = 0.0, 0.0
    for i, goal in enumerate(g["goals"]):
        p, f, it = so.optimize(goal)
        assert it < 100
        assert f <= g["res_f"][i] + 1e-9 * max(1.0, g["res_f"][i]), (i, f, g["res_f"][i])
        scale = np.array([1.0, 1.0, p[2]])
        worst_p = max(worst_p, float((np.abs(p - g["res_p"][i]) / scale).max()))
        worst_f = max(worst_f, float(g["res_f"][i] - f))
    print(f"restated solver vs scipy L-BFGS-B: worst parameter difference {worst_p:.2e}, largest objective gain {worst_f:.2e}")
    assert worst_p < 5e-4


def test_goal_state_set_vs_literal(golden):
    """``planner.goal_state_set`` is the reference's ``get_goal_state_set`` (local_planner.py:154-275), bit for bit."""
    from python_motionplanning_b200.planner import goal_state_set
    g = golden("spiral_opt.npz")
    wps = g["goalset_waypoints"].tolist()
    for row, want in zip(g["goalset_in"], g["goalset_out"]):
        gi = int(row[0])
        got = goal_state_set(gi, list(g["goalset_waypoints"][gi]), wps, list(row[1:4]) + [25.0])
        assert np.array_equal(np.array(got), want)
