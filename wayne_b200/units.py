"""Physical units for the host layer.

The reference talks astropy.units everywhere on the path
(wayne/exposure_generator.py:247-248, 275, 312, 489, 590-623;
wayne/detector.py:100, 246).  astropy is used when it is importable; otherwise
this module provides the small subset the path needs (a ``Quantity`` carrying a
numpy value and a ``Unit`` with an SI scale and dimension vector), so that
reference-style calling code -- ``1 * u.year``, ``scan_speed.to(u.pixel / u.ms)``,
``read_times.to(u.s).value`` -- keeps working on a machine without astropy.

The device engine never sees a Quantity: everything is reduced to plain floats
in documented units by :func:`value_in` at the API boundary.
"""
from __future__ import annotations

import numbers

import numpy as np

try:  # pragma: no cover - exercised only where astropy is installed
    import astropy.units as _apu

    HAVE_ASTROPY = True
except Exception:  # ImportError or a broken install
    _apu = None
    HAVE_ASTROPY = False


_DIMS = ("time", "length", "pixel", "count", "energy")


class UnitConversionError(ValueError):
    pass


class Unit(object):
    """scale * prod(base_i ** power_i); enough algebra for the simulator."""

    __array_priority__ = 1000

    def __init__(self, name, scale=1.0, dims=None):
        self.name = name
        self.scale = float(scale)
        self.dims = tuple(dims) if dims is not None else (0,) * len(_DIMS)

    # -- algebra ---------------------------------------------------------
    def _combine(self, other, sign):
        dims = tuple(a + sign * b for a, b in zip(self.dims, other.dims))
        if sign > 0:
            return Unit("{} {}".format(self.name, other.name), self.scale * other.scale, dims)
        return Unit("{} / ({})".format(self.name, other.name), self.scale / other.scale, dims)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, +1)
        if isinstance(other, Quantity):
            return Quantity(other.value, self * other.unit)
        return Quantity(other, self)

    def __rmul__(self, other):
        return Quantity(other, self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, -1)
        if isinstance(other, Quantity):
            return Quantity(1.0 / np.asarray(other.value), self / other.unit)
        return Quantity(1.0 / np.asarray(other, dtype=float), self)

    def __rtruediv__(self, other):
        inv = Unit("1 / ({})".format(self.name), 1.0 / self.scale, tuple(-d for d in self.dims))
        return Quantity(other, inv)

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __pow__(self, p):
        return Unit("({})**{}".format(self.name, p), self.scale ** p,
                    tuple(d * p for d in self.dims))

    def __eq__(self, other):
        return isinstance(other, Unit) and self.dims == other.dims and \
            np.isclose(self.scale, other.scale, rtol=1e-12)

    def __hash__(self):
        return hash((self.dims, round(self.scale, 12)))

    def __repr__(self):
        return "Unit({!r})".format(self.name)

    __str__ = lambda self: self.name

    def factor_to(self, other):
        """Number by which a value in ``self`` is multiplied to express it in ``other``."""
        if self.dims != other.dims:
            raise UnitConversionError("'{}' and '{}' are not convertible".format(self, other))
        return self.scale / other.scale

    def to(self, other, value=1.0):
        return value * self.factor_to(other)


def _dim(name):
    return tuple(1 if d == name else 0 for d in _DIMS)


class Quantity(object):
    __array_priority__ = 10000

    def __init__(self, value, unit):
        if isinstance(value, Quantity):
            value = value.to(unit).value
        self.value = value if isinstance(value, np.ndarray) else (
            np.asarray(value, dtype=float) if not isinstance(value, numbers.Number) else value)
        if isinstance(self.value, np.ndarray) and self.value.ndim == 0:
            self.value = self.value[()]
        self.unit = unit

    # -- conversion -------------------------------------------------------
    def to(self, unit):
        f = self.unit.factor_to(unit)
        return Quantity(self.value * f, unit)

    # -- container behaviour ----------------------------------------------
    def __len__(self):
        return len(self.value)

    def __getitem__(self, i):
        return Quantity(self.value[i], self.unit)

    def __iter__(self):
        for v in self.value:
            yield Quantity(v, self.unit)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def __float__(self):
        return float(self.value)

    def __bool__(self):
        return bool(np.any(self.value))

    __nonzero__ = __bool__

    def __repr__(self):
        return "<Quantity {} {}>".format(self.value, self.unit)

    # -- arithmetic --------------------------------------------------------
    def _other(self, other):
        if isinstance(other, Quantity):
            return other.to(self.unit).value
        if self.unit.dims == (0,) * len(_DIMS):
            return other
        raise UnitConversionError("cannot combine '{}' with a dimensionless value".format(self.unit))

    def __add__(self, other):
        return Quantity(self.value + self._other(other), self.unit)

    __radd__ = __add__

    def __sub__(self, other):
        return Quantity(self.value - self._other(other), self.unit)

    def __rsub__(self, other):
        return Quantity(self._other(other) - self.value, self.unit)

    def __neg__(self):
        return Quantity(-self.value, self.unit)

    def __mul__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.value * other.value, self.unit * other.unit)
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit * other)
        return Quantity(self.value * other, self.unit)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Quantity):
            return Quantity(self.value / other.value, self.unit / other.unit)
        if isinstance(other, Unit):
            return Quantity(self.value, self.unit / other)
        return Quantity(self.value / other, self.unit)

    def __rtruediv__(self, other):
        inv = Unit("1 / ({})".format(self.unit.name), 1.0 / self.unit.scale,
                   tuple(-d for d in self.unit.dims))
        return Quantity(other / self.value, inv)

    __div__ = __truediv__

    def _cmp(self, other, op):
        return op(self.value, self._other(other))

    def __eq__(self, other):
        try:
            return self._cmp(other, np.equal)
        except UnitConversionError:
            return False

    def __ne__(self, other):
        return np.logical_not(self.__eq__(other))

    def __lt__(self, other):
        return self._cmp(other, np.less)

    def __le__(self, other):
        return self._cmp(other, np.less_equal)

    def __gt__(self, other):
        return self._cmp(other, np.greater)

    def __ge__(self, other):
        return self._cmp(other, np.greater_equal)

    __hash__ = None


if HAVE_ASTROPY:  # pragma: no cover
    s, ms, minute, hour, day, year = _apu.s, _apu.ms, _apu.min, _apu.hour, _apu.day, _apu.year
    micron, angstrom, nm, m, cm = _apu.micron, _apu.angstrom, _apu.nm, _apu.m, _apu.cm
    pixel, count, ct, photon, ph, electron = (_apu.pixel, _apu.count, _apu.ct, _apu.photon,
                                              _apu.ph, _apu.electron)
    erg = _apu.erg
    dimensionless = _apu.dimensionless_unscaled
    Quantity = _apu.Quantity  # noqa: F811
else:
    dimensionless = Unit("", 1.0)
    s = Unit("s", 1.0, _dim("time"))
    ms = Unit("ms", 1e-3, _dim("time"))
    minute = Unit("min", 60.0, _dim("time"))
    hour = Unit("h", 3600.0, _dim("time"))
    day = Unit("d", 86400.0, _dim("time"))
    year = Unit("yr", 365.25 * 86400.0, _dim("time"))
    m = Unit("m", 1.0, _dim("length"))
    cm = Unit("cm", 1e-2, _dim("length"))
    micron = Unit("micron", 1e-6, _dim("length"))
    nm = Unit("nm", 1e-9, _dim("length"))
    angstrom = Unit("Angstrom", 1e-10, _dim("length"))
    pixel = Unit("pix", 1.0, _dim("pixel"))
    count = Unit("ct", 1.0, _dim("count"))
    ct = count
    photon = Unit("ph", 1.0, _dim("count"))
    ph = photon
    electron = Unit("electron", 1.0, _dim("count"))
    erg = Unit("erg", 1e-7, _dim("energy"))

# astropy spells it `min`
min = minute  # noqa: A001


def is_quantity(x):
    return hasattr(x, "unit") and hasattr(x, "value") and hasattr(x, "to")


def value_in(x, unit):
    """Plain float / ndarray of ``x`` expressed in ``unit``; bare numbers are
    taken to be in ``unit`` already (documented per argument at the call site)."""
    if is_quantity(x):
        return x.to(unit).value
    return x
