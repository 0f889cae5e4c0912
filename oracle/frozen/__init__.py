"""Frozen solver / kinematics copies used ONLY by the golden generator and the tests (see the module headers)."""
