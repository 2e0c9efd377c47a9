"""Unit handling at the Python boundary.

The reference passes astropy Quantities into ththmod (tau in us, fd/edges in
mHz, eta in s^3; scintools/ththmod.py:1639-1668).  astropy is optional here:
Quantities are accepted and converted when astropy is importable, bare numbers
are taken to be in the reference's default unit.  The kernels only ever see
float64 values in those default units.
"""
import numpy as np

try:  # pragma: no cover - astropy is not installed in the build image
    import astropy.units as _u
except Exception:  # noqa: BLE001
    _u = None

DEFAULTS = {"tau": "us", "fd": "mHz", "edges": "mHz", "eta": "s3",
            "time": "s", "freq": "MHz"}


def _ap(unit):
    if unit == "s3":
        return _u.s ** 3
    return getattr(_u, unit)


def value(x, unit):
    """Plain float64 value(s) of ``x`` expressed in ``unit`` ('us', 'mHz',
    's3', 's', 'MHz').  Incompatible astropy units raise UnitConversionError
    like the reference's unit_checks."""
    if _u is not None and isinstance(x, _u.Quantity):
        return np.asarray(x.to_value(_ap(unit)), dtype=np.float64)
    if hasattr(x, "to_value") and hasattr(x, "unit"):
        # duck-typed Quantity (e.g. the oracle's units shim in tests)
        import sys
        mod = sys.modules.get(type(x).__module__)
        name = {"s3": None}.get(unit, unit)
        tgt = (mod.s ** 3) if name is None else getattr(mod, name)
        return np.asarray(x.to_value(tgt), dtype=np.float64)
    return np.asarray(x, dtype=np.float64)


def wrap(val, unit, like=None):
    """Attach a unit on the way out when astropy is present and the caller
    used Quantities; otherwise return the bare value."""
    if _u is not None and (like is None or isinstance(like, _u.Quantity)):
        return val * _ap(unit)
    return val
