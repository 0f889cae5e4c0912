"""In-memory loader of the reference package (see oracle/refrun/__init__.py)."""
import importlib.machinery
import importlib.util
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("SIO_REFERENCE_ROOT", "/root/reference")
PKG = "semiinfinite_ref"
MODULES = ["utils", "objective", "constraint", "sip", "geometryopt", "planopt"]     # planopt: Klampt's planners are a stand-in hook (klampt_shim.PLANNER_FACTORY)


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "semiinfinite", "sip.py"))


_PRINT_STMT = re.compile(r'^(\s*)print\s+(?!\()(.+?)\s*$')


def py2_to_py3(src):
    """The mechanical fixes the reference needs; no logic is touched."""
    out = []
    for line in src.split("\n"):
        stripped = line.lstrip()
        if not stripped.startswith("#"):
            line = re.sub(r'\bxrange\(', 'range(', line)
            line = re.sub(r'\braw_input\(', 'input(', line)
            line = line.replace('.iteritems()', '.items()').replace('.itervalues()', '.values()').replace('.iterkeys()', '.keys()')
            if re.match(r'^from sip import \*\s*$', line):
                line = 'from .sip import *'
            m = _PRINT_STMT.match(line)
            if m and '"""' not in line:
                line = '%sprint(%s)' % (m.group(1), m.group(2))
        out.append(line)
    return "\n".join(out)


class _Proxy:
    """Attribute proxy: everything from `base` except the names overridden here."""

    def __init__(self, base, **over):
        self.__dict__['_base'] = base
        self.__dict__.update(over)

    def __getattr__(self, name):
        return getattr(self._base, name)


def _scipy_compat():
    """scipy of the reference's era (2018) let scipy.sparse.vstack/bmat take 1-D dense rows and
    coerced each block with coo_matrix(); scipy 1.18 refuses.  The reference stacks its constraint
    rows that way (sip.py:317), so its modules get a scipy whose vstack restores the coercion."""
    import scipy
    import scipy.io
    import scipy.sparse
    import scipy.sparse.linalg

    def vstack(blocks, format=None, dtype=None):
        blocks = [b if scipy.sparse.issparse(b) else scipy.sparse.coo_matrix(b) for b in blocks]
        return scipy.sparse.vstack(blocks, format=format, dtype=dtype)

    return _Proxy(scipy, sparse=_Proxy(scipy.sparse, vstack=vstack))


def load(force=False):
    """Returns the dict {module name: module} of the reference package, executing it on first use."""
    if not available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE_ROOT)
    if PKG in sys.modules and not force:
        return {m: sys.modules[PKG + "." + m] for m in MODULES}
    from . import klampt_shim, osqp_shim
    klampt_shim.install()
    osqp_shim.install()
    pkg = types.ModuleType(PKG)
    pkg.__path__ = []          # a package with no importable files: submodules are registered below
    pkg.__package__ = PKG
    sys.modules[PKG] = pkg
    mods = {}
    for name in MODULES:
        fn = os.path.join(REFERENCE_ROOT, "semiinfinite", name + ".py")
        src = py2_to_py3(open(fn).read())
        mod = types.ModuleType(PKG + "." + name)
        mod.__file__ = fn
        mod.__package__ = PKG
        sys.modules[PKG + "." + name] = mod
        setattr(pkg, name, mod)
        # docstrings of the reference hold py2 code with stray escapes: compile quietly
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            code = compile(src, fn, "exec")
        exec(code, mod.__dict__)
        if "scipy" in mod.__dict__:
            mod.__dict__["scipy"] = _scipy_compat()
        mods[name] = mod
    return mods
