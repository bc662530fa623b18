"""CPU oracle for the b200 motion-planning hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or as the timed CPU baseline.
The product package (``python_motionplanning_b200``) never imports this package and
raises loudly when its CUDA library is missing.

Contents
--------
``planar_numpy``     vectorised NumPy restatement of ``VehicleModel.planar_model`` /
                     ``planar_model_RK4`` (reference ``libs/vehicle_model/vehicle_model.py:220-445``)
``collision_numpy``  NumPy restatement of ``CollisionChecker.collision_check`` /
                     ``select_best_path_index`` (``libs/motionplanner/collision_checker.py:32-203``)
``lattice_numpy``    NumPy restatement of ``PathOptimizer.sample_spiral`` / ``transform_paths`` /
                     ``Env.box`` (input generators either side of the hot path)
``csrc/oracle.c``    the same algorithms in plain C (gcc, ``-ffp-contract=off``), used where the
                     NumPy oracle is too slow (full-size parity) and as the multi-threaded CPU baseline
``ref_loader``       imports the *unmodified* reference from ``/root/reference`` (exists only in the
                     build container) to pin the restatements and to generate ``tests/golden``
``make_golden``      the committed script that produced ``tests/golden/*.npz``

Pinning: the reference ships no tests and no golden vectors.  The oracle is pinned
by outputs of the literal reference executed in the build container (``tests/golden`` +
``make_golden.py``) and by the known-answer vectors of SURVEY.md Appendix C.
"""
