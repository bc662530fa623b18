"""Config 1 (BASELINE.json configs[0]): the reference's own single-vehicle closed loop, ``animate.py:27, 61-99`` ->
``libs/vehicle_model/drive.py:112-154``, unmodified.

* CPU: the harness on the literal reference reproduces ``tests/golden/closedloop_cfg1.npz`` (pins harness + staged copy).
* GPU: the same loop after ``install()`` -- every RK4 step, the collision fan-out and the path selection run on the
  engine -- agrees with the golden within 1e-9 / bit-exact flags and indices, for 25 frames (2,500 RK4 calls,
  25 planner calls) including the DataLog rows the loop writes.
"""
import json
import os

import numpy as np
import pytest

from oracle import config1_harness as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("no reference checkout: neither /root/reference nor baseline/_ref (run baseline/stage_reference.py)")
    return ref_loader


def test_staged_reference_is_verbatim():
    """baseline/_ref (what travels to the GPU box) is a byte-for-byte copy of the reference's files."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("stage_reference", os.path.join(ROOT, "baseline", "stage_reference.py"))
    st = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(st)
    if os.path.isdir("/root/reference"):
        assert st.stage() is not None
    if not os.path.isdir(st.DEST):
        pytest.skip("baseline/_ref not staged here")
    assert st.verify(), "baseline/_ref differs from its manifest"
    if os.path.isdir("/root/reference"):
        for rel, p in st._files("/root/reference"):
            assert open(p, "rb").read() == open(os.path.join(st.DEST, rel), "rb").read(), rel


def test_harness_reproduces_golden_on_literal_reference(golden):
    _need_reference()
    frames = 2
    res = H.run(frames, use_engine=False)
    assert res["kbm_class"].startswith("libs.vehicle_model")
    cmp_ = H.compare(res, golden("closedloop_cfg1.npz"), frames, tol=1e-12)
    assert cmp_["ok"], cmp_


@pytest.mark.gpu
def test_config1_car_drive_runs_on_engine(golden, engine):
    _need_reference()
    frames = 25
    res = H.run(frames, use_engine=True)
    assert res["kbm_class"] == "python_motionplanning_b200.vehicle_model.VehicleModel"
    cmp_ = H.compare(res, golden("closedloop_cfg1.npz"), frames, tol=1e-9)
    cmp_.update(frames=frames, rk4_calls=frames * 100, planner_calls=frames,
                s_per_frame_median=float(np.median(res["frame_s"])), s_per_frame_first=float(res["frame_s"][0]),
                rebound=res["rebound"])
    print("config 1 on the engine:", json.dumps(cmp_))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "config1_closed_loop.json"), "w") as f:
            json.dump(cmp_, f, indent=1)
    assert cmp_["ok"], cmp_
