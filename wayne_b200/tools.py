"""Array helpers that sit on the exposure-synthesis path.

Own implementations of the three helpers the path uses (scope table a6, a9):
``crop_spectrum_ind`` (wayne/tools.py:46-77), ``bin_centers_to_widths``
(:106-128), ``crop_central_box`` (:317-324) plus ``bin_centers_to_edges`` and
``crop_spectrum`` which the reference's own tests pin (tests/test_tools.py).
"""
import numpy as np

from . import units as u


def _plain(x):
    return np.asarray(u.value_in(x, getattr(x, "unit", None)) if u.is_quantity(x) else x, dtype=float)


def crop_spectrum_ind(min_wl, max_wl, wl):
    """Slice bounds (imin, imax) of the part of an ORDERED ``wl`` inside [min_wl, max_wl].

    Same selection rule as the reference: imin is the element closest above (or
    at) ``min_wl``; imax-1 the element closest below (or at) ``max_wl``.
    """
    if u.is_quantity(wl):
        unit = wl.unit
        w = np.asarray(wl.value, dtype=float)
        lo = u.value_in(min_wl, unit)
        hi = u.value_in(max_wl, unit)
    else:
        w = np.asarray(wl, dtype=float)
        lo, hi = float(u.value_in(min_wl, u.micron)), float(u.value_in(max_wl, u.micron))
    above = w - lo
    above = np.where(above < 0, above.max(), above)   # entries below the limit can never win
    imin = int(above.argmin())
    below = w - hi
    below = np.where(below > 0, below.min(), below)
    imax = int(below.argmax()) + 1
    return imin, imax


def crop_spectrum(min_wl, max_wl, wl, flux):
    imin, imax = crop_spectrum_ind(min_wl, max_wl, wl)
    return wl[imin:imax], flux[imin:imax]


def _half_gaps(centers):
    c = centers
    gaps = (c - np.roll(c, 1)) / 2.
    gaps[0] = gaps[1]          # the first bin borrows the gap to its right-hand neighbour
    return gaps


def bin_centers_to_edges(centers):
    """Edges half-way between neighbouring centres; the outer edges mirror the
    adjacent half-gap (tests/test_tools.py:55-62)."""
    c = np.asarray(centers, dtype=float)
    g = _half_gaps(c)
    edges = np.empty(len(c) + 1)
    edges[:-1] = c - g
    edges[-1] = c[-1] + g[-1]
    return edges


def bin_centers_to_widths(centers):
    """Width of each bin = half-gap to the previous centre + half-gap to the next
    (edge bins count their single neighbour's half-gap twice).
    KAT: [1, 2, 4, 5.4] -> [1, 1.5, 1.7, 1.4] (tests/test_tools.py:75-81)."""
    unit = None
    if u.is_quantity(centers):
        unit = centers.unit
        centers = centers.value
    c = np.asarray(centers, dtype=float)
    g = _half_gaps(c)
    nxt = np.roll(g, -1)
    nxt[-1] = g[-1]
    w = g + nxt
    return w * unit if unit is not None else w


def crop_central_box(array, size):
    """Central ``size`` x ``size`` box of a square array.

    The reference slices ``array[i:-i]`` with ``i = (len - size) / 2`` which is
    EMPTY for ``size >= len`` (every 1024 full-frame call
    with sky / gain / flat enabled raises there).  Defined behaviour here:
    identity when ``size >= len(array)``.
    """
    n = len(array)
    if size >= n:
        return array
    i = (n - size) // 2
    return array[i:n - i, i:n - i]
