"""Minimal arithmetic stand-in for ``astropy.units`` (astropy is not installable
offline) so that the reference's own ``ththmod`` source can be EXECUTED in the
build container by ``oracle/ref_loader.py`` / ``oracle/make_golden.py``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Written from the documented
behaviour of astropy Quantities, no astropy code was available to consult:

* a unit is (scale, {base: power}); conversion factor a->b is
  ``a.scale / b.scale`` and a factor that is exactly 1.0 applies no arithmetic;
* add / subtract / compare / floor_divide convert the 2nd operand to the unit
  of the 1st; multiply / divide / power combine units; results of comparisons
  are plain bool arrays;
* indexing, ``.T``, reductions (mean / max / min / sum) keep the unit.

With us = 1e-6 s, mHz = 1e-3 / s: s^3 mHz^2 -> us has factor
(1e-3)**2 / 1e-6 == 1.0 exactly, 1/s -> mHz is 1000.0 and 1/MHz -> us is 1.0.
Only the operations the theta-theta hot path uses are implemented.
"""
import numpy as np

_BASES = ("s", "m")


class UnitConversionError(ValueError):
    pass


class UnitBase:
    __array_priority__ = 100000
    __array_ufunc__ = None   # make ndarray (in-place) binops defer to us

    def __init__(self, scale, powers, name=None):
        self.scale = float(scale)
        self.powers = {k: v for k, v in powers.items() if v != 0}
        self.name = name

    # ---- algebra -------------------------------------------------------
    def _combine(self, other, sign):
        p = dict(self.powers)
        for k, v in other.powers.items():
            p[k] = p.get(k, 0) + sign * v
        sc = self.scale * other.scale if sign > 0 else self.scale / other.scale
        return UnitBase(sc, p)

    def __mul__(self, other):
        if isinstance(other, UnitBase):
            return self._combine(other, +1)
        return Quantity(other, self)

    def __rmul__(self, other):
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, UnitBase):
            return self._combine(other, -1)
        return Quantity(1.0 / np.asarray(other), self)

    def __rtruediv__(self, other):
        inv = UnitBase(1.0 / self.scale,
                       {k: -v for k, v in self.powers.items()})
        return Quantity(other, inv)

    def __pow__(self, p):
        return UnitBase(self.scale ** p,
                        {k: v * p for k, v in self.powers.items()})

    # ---- comparison / conversion --------------------------------------
    def is_equivalent(self, other):
        return self.powers == other.powers

    def _to(self, other):
        if self is other:
            return 1.0
        if not self.is_equivalent(other):
            raise UnitConversionError("%r -> %r" % (self, other))
        return self.scale / other.scale

    def __eq__(self, other):
        return isinstance(other, UnitBase) and self.powers == other.powers \
            and self.scale == other.scale

    def __hash__(self):
        return hash((self.scale, tuple(sorted(self.powers.items()))))

    def __repr__(self):
        return self.name or "Unit(%g, %r)" % (self.scale, self.powers)

    __str__ = __repr__


Unit = UnitBase
dimensionless_unscaled = UnitBase(1.0, {}, "")
s = UnitBase(1.0, {"s": 1}, "s")
us = UnitBase(1e-6, {"s": 1}, "us")
ms = UnitBase(1e-3, {"s": 1}, "ms")
min = UnitBase(60.0, {"s": 1}, "min")  # noqa: A001  (astropy name)
Hz = UnitBase(1.0, {"s": -1}, "Hz")
mHz = UnitBase(1e-3, {"s": -1}, "mHz")
MHz = UnitBase(1e6, {"s": -1}, "MHz")
m = UnitBase(1.0, {"m": 1}, "m")
km = UnitBase(1e3, {"m": 1}, "km")


def _unit_of(x):
    return x.unit if isinstance(x, Quantity) else dimensionless_unscaled


def _val(x):
    return x.view(np.ndarray) if isinstance(x, Quantity) else np.asarray(x)


def _convert(x, unit):
    """Value of x expressed in ``unit`` (no arithmetic when factor == 1)."""
    # astropy lets bare 0 / inf / nan stand in for any unit (``tau < 0``)
    if not isinstance(x, (Quantity, UnitBase)):
        a = np.asarray(x)
        if a.dtype.kind in "fiub" and a.size and bool(np.all((a == 0) | ~np.isfinite(a))):
            return a
    f = _unit_of(x)._to(unit)
    v = _val(x)
    return v if f == 1.0 else v * f


_SAME_UNIT_BINARY = {np.add, np.subtract, np.maximum, np.minimum, np.fmod,
                     np.remainder, np.mod}
_COMPARE = {np.less, np.less_equal, np.greater, np.greater_equal, np.equal,
            np.not_equal}
_KEEP_UNIT_UNARY = {np.absolute, np.negative, np.positive, np.fabs,
                    np.conjugate, np.rint, np.floor, np.ceil}


class Quantity(np.ndarray):
    __array_priority__ = 10000

    def __new__(cls, value, unit=dimensionless_unscaled, copy=True):
        if isinstance(value, Quantity):
            unit = value.unit * unit if unit is not dimensionless_unscaled \
                else value.unit
            value = value.view(np.ndarray)
        arr = np.array(value) if copy else np.asarray(value)
        if arr.dtype.kind in "iu" or arr.dtype.kind == "b":
            arr = arr.astype(float)
        obj = arr.view(cls)
        obj.unit = unit
        return obj

    def __array_finalize__(self, obj):
        self.unit = getattr(obj, "unit", dimensionless_unscaled)

    # ---- accessors -----------------------------------------------------
    @property
    def value(self):
        v = self.view(np.ndarray)
        return v[()] if v.ndim == 0 else v

    def to(self, unit):
        f = self.unit._to(unit)
        v = self.view(np.ndarray)
        return Quantity(v if f == 1.0 else v * f, unit)

    def to_value(self, unit=None):
        if unit is None:
            return self.value
        f = self.unit._to(unit)
        v = self.view(np.ndarray)
        out = v if f == 1.0 else v * f
        return out[()] if out.ndim == 0 else out

    def astype(self, dtype, **kw):
        out = self.view(np.ndarray).astype(dtype, **kw).view(Quantity)
        out.unit = self.unit
        return out

    def __getitem__(self, key):
        out = self.view(np.ndarray)[key]
        q = np.asarray(out).view(Quantity)
        q.unit = self.unit
        return q

    def __bool__(self):
        return bool(self.view(np.ndarray))

    def __float__(self):
        return float(self.view(np.ndarray))

    def __int__(self):
        return int(self.view(np.ndarray))

    def __index__(self):
        return int(self.view(np.ndarray))

    def __repr__(self):
        return "<Quantity %r %s>" % (self.view(np.ndarray), self.unit)

    __str__ = __repr__

    def __format__(self, spec):
        return "%s %s" % (format(self.value, spec), self.unit)

    # ---- unit operands -------------------------------------------------
    def __mul__(self, other):
        if isinstance(other, UnitBase):
            return Quantity(self.view(np.ndarray), self.unit * other,
                            copy=False)
        return np.multiply(self, other)

    __rmul__ = __mul__

    def __imul__(self, other):
        if isinstance(other, UnitBase):
            self.unit = self.unit * other
            return self
        return np.multiply(self, other)

    def __truediv__(self, other):
        if isinstance(other, UnitBase):
            return Quantity(self.view(np.ndarray), self.unit / other,
                            copy=False)
        return np.true_divide(self, other)

    def __itruediv__(self, other):
        return self.__truediv__(other)

    def __isub__(self, other):
        return np.subtract(self, other)

    def __iadd__(self, other):
        return np.add(self, other)

    # ---- ufuncs --------------------------------------------------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        out = kwargs.pop("out", None)
        if out is not None:
            kwargs.pop("subok", None)
            kwargs.pop("casting", None)
            res = self.__array_ufunc__(ufunc, method, *inputs, **kwargs)
            tgt = out[0]
            np.copyto(_val(tgt), _val(res), casting="unsafe")
            if isinstance(tgt, Quantity):
                tgt.unit = _unit_of(res)
            return tgt
        if method == "reduce":
            res = getattr(ufunc, method)(_val(inputs[0]), **kwargs)
            if ufunc in (np.add, np.maximum, np.minimum):
                return Quantity(res, _unit_of(inputs[0]), copy=False)
            return res
        if method != "__call__":
            return NotImplemented
        if ufunc in _SAME_UNIT_BINARY:
            a, b = inputs
            ua = _unit_of(a)
            if not isinstance(a, Quantity) and not ua.is_equivalent(_unit_of(b)):
                ua = _unit_of(b)
            res = ufunc(_convert(a, ua), _convert(b, ua), **kwargs)
            return Quantity(res, ua, copy=False)
        if ufunc in _COMPARE:
            a, b = inputs
            ua = _unit_of(a) if isinstance(a, Quantity) else _unit_of(b)
            return ufunc(_convert(a, ua), _convert(b, ua), **kwargs)
        if ufunc is np.floor_divide:
            a, b = inputs
            ua = _unit_of(a)
            res = ufunc(_convert(a, ua), _convert(b, ua), **kwargs)
            return Quantity(res, dimensionless_unscaled, copy=False)
        if ufunc is np.multiply:
            a, b = inputs
            res = ufunc(_val(a), _val(b), **kwargs)
            return Quantity(res, _unit_of(a) * _unit_of(b), copy=False)
        if ufunc in (np.true_divide, np.divide):
            a, b = inputs
            res = ufunc(_val(a), _val(b), **kwargs)
            return Quantity(res, _unit_of(a) / _unit_of(b), copy=False)
        if ufunc is np.power:
            a, p = inputs
            res = ufunc(_val(a), _val(p), **kwargs)
            return Quantity(res, _unit_of(a) ** float(p), copy=False)
        if ufunc is np.square:
            res = ufunc(_val(inputs[0]), **kwargs)
            return Quantity(res, _unit_of(inputs[0]) ** 2, copy=False)
        if ufunc is np.sqrt:
            res = ufunc(_val(inputs[0]), **kwargs)
            return Quantity(res, _unit_of(inputs[0]) ** 0.5, copy=False)
        if ufunc in _KEEP_UNIT_UNARY:
            res = ufunc(_val(inputs[0]), **kwargs)
            return Quantity(res, _unit_of(inputs[0]), copy=False)
        if ufunc in (np.isfinite, np.isnan, np.isinf, np.sign, np.signbit):
            return ufunc(_val(inputs[0]), **kwargs)
        # anything else: dimensionless only
        vals = []
        for x in inputs:
            if isinstance(x, Quantity) and x.unit.powers:
                raise UnitConversionError(
                    "%s needs dimensionless input" % ufunc.__name__)
            vals.append(_convert(x, dimensionless_unscaled))
        return ufunc(*vals, **kwargs)
