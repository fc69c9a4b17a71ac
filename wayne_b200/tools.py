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


# ---------------------------------------------------------------------------
# spectrum preparation for the visit driver (upstream of the
# exposure path, host-side, once per visit)
# ---------------------------------------------------------------------------
def load_and_sort_spectrum(file_path):
    """Two whitespace-separated columns (wl, flux or depth), sorted by wl
    (wayne/tools.py:172-186)."""
    data = np.loadtxt(file_path, dtype=np.float64, ndmin=2)
    order = np.argsort(data[:, 0], kind='stable')
    return data[order, 0], data[order, 1]


def order_flux_grid(wavelength, spectrum):
    order = np.argsort(np.asarray(wavelength), kind='stable')
    return np.asarray(wavelength, dtype=float)[order], np.asarray(spectrum, dtype=float)[order]


def load_pheonix_stellar_grid_fits(fits_file):
    """PHOENIX model table (columns Wavelength, Flux), sorted, duplicate
    wavelengths removed (wayne/tools.py:152-170; spelling kept)."""
    from . import fitsio
    with fitsio.open(fits_file) as f:
        tab = f[1].data
        wl, flux = order_flux_grid(tab['Wavelength'], tab['Flux'])
    keep = np.nonzero(np.diff(wl))
    return wl[keep], flux[keep]


def wl_at_resolution(R, wl_min, wl_max):
    """Evenly spaced grid with spacing (mid wavelength) / R (wayne/tools.py:303-314)."""
    mid_wl = (wl_max - wl_min) / 2 + wl_min
    delta_wl = mid_wl / R
    return np.arange(wl_min, wl_max + delta_wl, delta_wl)


def rebin_spec(wavelength, spectrum, new_wavelength):
    """Flux-conserving rebinning onto the bins centred on ``new_wavelength``.

    The reference delegates to pysynphot (wayne/tools.py:131-149), which is not
    available; this is the same definition -- the mean of the piecewise-linear
    input spectrum over each output bin (bin edges half-way between centres) --
    written out: the cumulative trapezoid integral of the input is interpolated
    at the bin edges and differenced."""
    wl = np.asarray(wavelength, dtype=float)
    sp = np.asarray(spectrum, dtype=float)
    new = np.asarray(new_wavelength, dtype=float)
    edges = bin_centers_to_edges(new)
    cum = np.concatenate([[0.0], np.cumsum(0.5 * (sp[1:] + sp[:-1]) * np.diff(wl))])
    # integral up to an arbitrary x: cum at the left node + trapezoid of the partial segment
    xe = np.clip(edges, wl[0], wl[-1])
    k = np.clip(np.searchsorted(wl, xe, side='right') - 1, 0, len(wl) - 2)
    dx = xe - wl[k]
    slope = (sp[k + 1] - sp[k]) / (wl[k + 1] - wl[k])
    integ = cum[k] + sp[k] * dx + 0.5 * slope * dx * dx
    width = np.diff(xe)
    with np.errstate(invalid='ignore', divide='ignore'):
        out = np.where(width > 0, np.diff(integ) / width, 0.0)
    return out


def blackbody_lambda(wl_micron, temperature):
    """Planck B_lambda [erg / (s cm^2 angstrom sr)] at wl [micron] (stands in for
    astropy's blackbody_lambda, wayne/run_visit.py:203)."""
    h, c, k = 6.62607015e-27, 2.99792458e10, 1.380649e-16
    lam = np.asarray(wl_micron, dtype=float) * 1e-4            # cm
    b = 2 * h * c * c / lam ** 5 / np.expm1(h * c / (lam * k * temperature))   # per cm
    return b * 1e-8                                            # per angstrom
