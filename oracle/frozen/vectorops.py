"""FROZEN COPY (oracle/frozen, round 2) of semiinfiniteoptimization_b200/_host/vectorops.py as of commit 92f6a08 -- TEST INFRASTRUCTURE.

The golden generator (tools/make_golden.py -> oracle/refrun) answers the reference's Klampt / OSQP calls with THIS file,
not with the product's module, so a change to the product can no longer rewrite tests/golden/.  Do not edit it to follow
the product: the product is tested AGAINST it (tests/test_frozen_cpu.py) at a stated tolerance.

Original module docstring follows.

Minimal list-based vector helpers with the names the reference's call sites use
(`klampt.math.vectorops`; call sites: geometryopt.py:161-168, 260-279, 384, 391, 415-417)."""
import math


def add(a, b):
    return [x + y for x, y in zip(a, b)]


def sub(a, b):
    return [x - y for x, y in zip(a, b)]


def mul(a, s):
    return [x * s for x in a]


def div(a, s):
    return [x / s for x in a]


def madd(a, b, c):
    return [x + c * y for x, y in zip(a, b)]


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def normSquared(a):
    return sum(x * x for x in a)


def norm(a):
    return math.sqrt(normSquared(a))


def distance(a, b):
    return norm(sub(a, b))


def cross(a, b):
    return [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]


def unit(a):
    n = norm(a)
    return list(a) if n == 0 else div(a, n)


def interpolate(a, b, u):
    return [x + u * (y - x) for x, y in zip(a, b)]
