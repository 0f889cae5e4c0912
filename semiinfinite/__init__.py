"""Drop-in import path: code written against the reference does `from semiinfinite.geometryopt import *`
/ `from semiinfinite.sip import ...` (geomopt.py:10-11, robotposeopt.py:10-11).  These modules forward
to the B200-native implementation in `semiinfiniteoptimization_b200`."""
