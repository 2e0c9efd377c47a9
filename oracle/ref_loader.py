"""Import the UNMODIFIED reference (/root/reference/scintools) in the build
container, where its optional dependencies are missing.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable where
/root/reference exists (the build container); the GPU box never has it, so
nothing in tests marked ``gpu``, ``smoke()`` or ``bench.py`` calls this module.
``oracle/make_golden.py`` uses it to generate ``tests/golden/*.npz``.

matplotlib / lmfit / skimage / emcee ... are replaced by MagicMock modules
(never executed on the paths we run).  ``astropy.units`` is replaced by
``oracle.units_shim`` (a small arithmetic stand-in) so that ``ththmod`` can run
too; ``astropy.constants/time/coordinates/io`` are mocks.
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("SCINTOOLS_REFERENCE", "/root/reference")

_MOCKS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
    "matplotlib.gridspec", "matplotlib.dates", "matplotlib.ticker",
    "lmfit", "skimage", "skimage.restoration", "emcee", "bilby", "corner",
    "astropy.constants", "astropy.time", "astropy.coordinates", "astropy.io",
    "astropy.io.fits", "astropy.utils", "astropy.utils.iers",
]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scintools"))


def load():
    """Return the reference ``scintools`` package (cached in sys.modules)."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    if "scintools" in sys.modules and getattr(
            sys.modules["scintools"], "_b200_oracle_loaded", False):
        return sys.modules["scintools"]
    try:
        import astropy  # noqa: F401  (real astropy wins if it ever exists)
    except Exception:
        from oracle import units_shim
        ap = types.ModuleType("astropy")
        ap.units = units_shim
        ap.__path__ = []
        sys.modules["astropy"] = ap
        sys.modules["astropy.units"] = units_shim
    for name in _MOCKS:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = MagicMock(name=name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the package __init__ star-imports everything; import the modules we need
    pkg = types.ModuleType("scintools")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "scintools")]
    pkg._b200_oracle_loaded = True
    sys.modules["scintools"] = pkg
    for sub in ("scint_utils", "scint_models", "scint_sim", "ththmod",
                "dynspec"):
        mod = importlib.import_module("scintools." + sub)
        setattr(pkg, sub, mod)
    return pkg
