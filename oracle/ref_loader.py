"""
TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules.

Loads ``/root/reference/fastbox/{box,beams,halos,tracers}.py`` by file path,
with stub ``pyccl`` / ``pylab`` modules in ``sys.modules`` (the reference's
``fastbox/__init__.py:2-3`` eagerly imports modules whose dependencies are not
installed, so ``import fastbox`` cannot be used).  The stub ``pyccl`` forwards
to ``fastbox_b200.cosmology`` so oracle and product consume the same P(k).

Only available in the build container (``/root/reference`` does not exist on
the GPU box).  Used by ``oracle/make_golden.py`` to generate the fixtures in
``tests/golden/`` and by CPU tests (skipped when the reference is absent) to
pin ``oracle/restate.py`` against the real reference code.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FASTBOX_REFERENCE", "/root/reference")
_cache = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "fastbox", "box.py"))


def _install_stubs():
    import scipy.integrate
    if not hasattr(scipy.integrate, "simps"):       # removed in SciPy >= 1.14; box.py:680,892,899
        scipy.integrate.simps = scipy.integrate.simpson
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    if root not in sys.path:
        sys.path.insert(0, root)
    from fastbox_b200 import cosmology as cos
    if "pyccl" not in sys.modules or getattr(sys.modules["pyccl"], "_fb_stub", False):
        ccl = types.ModuleType("pyccl")
        ccl._fb_stub = True
        for name in ("Cosmology", "linear_matter_power", "nonlin_matter_power",
                     "h_over_h0", "growth_rate", "growth_factor",
                     "comoving_angular_distance"):
            setattr(ccl, name, getattr(cos, name))
        sys.modules["pyccl"] = ccl
    if "pylab" not in sys.modules:
        sys.modules["pylab"] = types.ModuleType("pylab")


def load(name):
    """Return reference module ``fastbox/<name>.py`` executed from its own source."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    path = os.path.join(REFERENCE_ROOT, "fastbox", name + ".py")
    spec = importlib.util.spec_from_file_location("fastbox_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def load_pkg(name):
    """
    Reference module that uses package-relative imports (``filters.py`` imports ``.foregrounds``):
    imported as ``fastbox_refpkg.<name>`` from a synthetic package whose path is the reference's
    ``fastbox/`` directory, so the reference's own ``__init__`` (which pulls in absent dependencies)
    is not executed and the source files stay unmodified.
    """
    key = "pkg:" + name
    if key in _cache:
        return _cache[key]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    if "fastbox_refpkg" not in sys.modules:
        pkg = types.ModuleType("fastbox_refpkg")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "fastbox")]
        sys.modules["fastbox_refpkg"] = pkg
    mod = importlib.import_module("fastbox_refpkg." + name)
    _cache[key] = mod
    return mod

