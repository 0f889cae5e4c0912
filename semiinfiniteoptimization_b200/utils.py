"""The reference's gradient-checking aids (semiinfinite/utils.py:3-51), same names and arguments.  Debugging helpers of the
SIP loop (sip.py:563, 931 call test_gradient when its gradient-test flag is on), not part of the hot path."""
import numpy as np


def numeric_gradient(f, x, h=1e-4, method='centered'):
    """Finite-difference gradient of the scalar function f at the vector x; h is the step.
    method: 'centered' (default), 'forward' or 'backward'.  In the one-sided methods the reference does not put a
    component back before moving on to the next, so component i is taken at x + h (e_0 + ... + e_i)   # quirk, reproduced"""
    xtest = np.array(x, dtype=np.float64)
    g = np.zeros(len(x))
    if method == 'centered':
        for i in range(len(x)):
            xtest[i] += h
            b = f(xtest)
            xtest[i] -= 2 * h
            a = f(xtest)
            xtest[i] = x[i]
            g[i] = b - a
        g *= 1.0 / (2 * h)
        return g
    if method == 'forward':
        a = f(x)
        for i in range(len(x)):
            xtest[i] += h
            g[i] = f(xtest) - a
        g *= 1.0 / h
        return g
    if method == 'backward':
        a = f(x)
        for i in range(len(x)):
            xtest[i] -= h
            g[i] = a - f(xtest)
        g *= 1.0 / h
        return g
    raise ValueError("Invalid method %s" % (method,))


def test_gradient(f, fgrad, x, h=1e-4, epsilon=1e-3, name=""):
    """True iff fgrad(x) agrees with the centred-difference gradient of f at x to within epsilon in every component;
    the components that do not are printed."""
    g_numeric = numeric_gradient(f, x, h)
    g_calc = fgrad(x)
    valid = True
    for i, v in enumerate(g_numeric - g_calc):
        if abs(v) > epsilon:
            print("test_gradient %s: errorneous value %d: numeric %g vs function %g" % (name, i, g_numeric[i], g_calc[i]))
            valid = False
    return valid


test_gradient.__test__ = False          # a helper with a pytest-looking name, not a test
