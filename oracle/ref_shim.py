"""Load the two reference files that still import under Python 3, unmodified, from /root/reference.

Test infrastructure only.  Used by ``oracle/make_golden.py`` (fixture generation) and by
``tests/test_oracle_vs_reference.py`` (skipped when /root/reference is absent, e.g. on the GPU box).

What is stubbed, and why that is semantically transparent:
  * ``spuq`` package __init__ (imports scipy.special.btdtri, gone in scipy 1.18) -> empty package
  * ``spuq.utils.type_check.takes`` -> identity decorator (argument type checks only)
  * ``spuq.utils.decorators.simple_int_cache`` -> identity (memoisation only)
  * ``xrange`` -> range, ``scipy.factorial`` -> scipy.special.factorial
Nothing is copied: the modules are executed from where they lie.
"""
import builtins
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("SPUQ_REFERENCE", "/root/reference")
_SRC = os.path.join(REF_ROOT, "src", "spuq")


def available():
    return os.path.isfile(os.path.join(_SRC, "math_utils", "multiindex_set.py"))


def _stub_package(name, path):
    mod = types.ModuleType(name)
    mod.__path__ = [path]
    sys.modules[name] = mod
    return mod


def load():
    """Return (multiindex_set module, polynomials module) of the reference."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    if "spuq.polyquad.polynomials" in sys.modules and getattr(
            sys.modules.get("spuq"), "_oracle_shim", False):
        return (sys.modules["spuq.math_utils.multiindex_set"],
                sys.modules["spuq.polyquad.polynomials"])

    import scipy
    import scipy.special
    if not hasattr(builtins, "xrange"):
        builtins.xrange = range
    if not hasattr(scipy, "factorial"):
        scipy.factorial = scipy.special.factorial

    pkg = _stub_package("spuq", _SRC)
    pkg._oracle_shim = True
    _stub_package("spuq.math_utils", os.path.join(_SRC, "math_utils"))
    _stub_package("spuq.polyquad", os.path.join(_SRC, "polyquad"))
    utils = _stub_package("spuq.utils", os.path.join(_SRC, "utils"))

    tc = types.ModuleType("spuq.utils.type_check")
    tc.takes = lambda *a, **k: (lambda f: f)
    tc.returns = lambda *a, **k: (lambda f: f)
    tc.anything = object
    tc.optional = lambda *a: object
    sys.modules["spuq.utils.type_check"] = tc
    utils.type_check = tc

    dec = types.ModuleType("spuq.utils.decorators")
    dec.simple_int_cache = lambda size: (lambda f: f)
    sys.modules["spuq.utils.decorators"] = dec
    utils.decorators = dec

    mis = importlib.import_module("spuq.math_utils.multiindex_set")
    importlib.import_module("spuq.polyquad._polynomials")
    pol = importlib.import_module("spuq.polyquad.polynomials")
    return mis, pol
