"""CPU oracle for the GPS/SLAM fusion hot path -- TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the arithmetic in the reference's
``EKFGPSSLAM.py`` (A2ureeE/GPS-optimize-SLAM).  It is the checker for the CUDA
path, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``gps_optimize_slam_b200/`` imports it, and the product raises
if its CUDA library is missing instead of falling back to this code.

Pinning status (see DESIGN.md "Oracle"):
  * fusion_oracle.py  -- PINNED: checked function-by-function against the
    unmodified reference imported in the build container (oracle/ref_loader.py)
    on both shipped fixture pairs and seeded synthetic trajectories; the
    resulting vectors are committed under tests/golden/ (oracle/make_golden.py).
  * utm_kruger.py     -- PARITY UNPINNED: the reference gets its UTM
    projection from pyproj/PROJ, which is not installed here and which the
    reference does not vendor or version-pin; there are no reference test
    vectors for it.  The restatement follows the published Karney/Krueger
    6th-order series (the algorithm PROJ documents for +proj=utm) and is
    checked against closed-form known answers only.
"""
