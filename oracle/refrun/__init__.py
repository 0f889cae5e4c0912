"""TEST INFRASTRUCTURE -- runs the UNMODIFIED reference Python package under this interpreter.

The reference (`/root/reference/semiinfinite/*.py`) is Python 2 on top of Klampt and OSQP, neither of
which exists in this image.  This package makes its own code runnable anyway, so that the golden
vectors under tests/golden/ come from the reference's OWN sip.py / geometryopt.py / constraint.py /
objective.py logic rather than from a restatement:

  load.py        reads the reference sources where they lie (never copied into this repository),
                 applies the mechanical py2 -> py3 source fixes in memory (xrange, iteritems,
                 raw_input, one print statement, one implicit relative import) and executes them as
                 the package `semiinfinite_ref`
  klampt_shim.py stand-in `klampt` modules: geometry containers from the host substrate (_host/geometry.py),
                 kinematics / so3 / se3 / trajectories from the frozen copies (oracle/frozen/), the geometric queries (distance_point / distance_ext) are the CPU oracle
                 (oracle/oracle.py) -- so the Klampt-side arithmetic is still a
                 restatement ("parity unpinned" for that layer), while everything above it is pinned
  osqp_shim.py   stand-in `osqp` module on the frozen ADMM solver (oracle/frozen/qp.py)

Only tools/make_golden.py and tests/ import this; it needs /root/reference and therefore only runs
in the build container, never on the GPU box.
"""
