"""WFC3 IR grisms G141 / G102: trace, dispersion, PSF, sensitivity, flat, sky.

Host-side mirror of ``wayne.grism`` (wayne/grism.py:24-817) keeping the public
names the path and its callers use (``G141``, ``G102``, ``get_trace``,
``wl_limits``, ``set_current_wavelength_only_dependent_array``, ``current_*``,
``get_flat_field``, ``get_master_sky``, ``_SpectrumTrace`` and the coefficient
tuples).  The heavy evaluation of these models -- per wavelength bin and per
sub-sample -- is done by the CUDA kernels in csrc/stage1.cuh and
csrc/gather.cuh; the numpy methods here give callers the same answers on the
host and are what tests/test_grism.py (lifted from the reference's KATs) pins.
Plotting helpers of the reference are not part of the path and are not kept.
"""
from __future__ import annotations

import numpy as np

from . import fitsio as fits
from . import params, tools
from . import units as u
from .detector import WFC3_IR

# aXe trace / wavelength-solution coefficients a0..a8, b0..b8 (Kuntschner et al.
# 2009 WFC3 ISRs, as tabulated in wayne/grism.py:756-776).
g141_trace_coeff = (1.96882, 9.09159E-5, -1.93260E-3, 1.04275E-2, -7.96978E-6,
                    -2.49607E-6, 1.45963E-9, 1.39757E-8, 4.8494E-10)
g141_wl_solution = (8.95431E3, 9.35925E-2, 0, 4.51423E1, 3.17239E-4,
                    2.17055E-3, -7.42504E-7, 3.48639E-7, 3.09213E-7)
g102_trace_coeff = (-3.55018E-1, 3.28722E-5, -1.44571E-3, 1.42852E-2,
                    -7.20713E-6, -2.42542E-6, 1.18294E-9, 1.19634E-8,
                    6.17274E-10)
g102_wl_solution = (6.38738E3, 4.55507E-2, 0, 2.35716E1, 3.60396E-4,
                    1.58739E-3, -4.25234E-7, -6.53726E-8, 0.)

# double-Gaussian PSF polynomials in wavelength [micron], highest power first
PSF_RATIO_POLY = (-0.25063428, 0.8332488, -0.80546074, 0.39896516)
PSF_SIGMAL_POLY = (0.69245668, -2.1043046, 2.22284446, -0.29689335)
PSF_SIGMAH_POLY = (2.90366189, -8.81859432, 8.96049229, 2.254503)

ANGSTROM_TO_MICRON = 1e-4


def wavelength_calibration_coeffs(x_ref, y_ref, trace_coeff, wl_sol_coeff):
    """Field-dependent trace slope/offset and dispersion slope/offset
    (m_t, c_t, m_w, c_w) at the source position."""
    a, b = trace_coeff, wl_sol_coeff
    xx, yy = x_ref ** 2, y_ref ** 2
    m_t = np.array(a[3] + a[4] * x_ref + a[5] * y_ref + a[6] * xx + a[7] * x_ref * y_ref + a[8] * yy)
    c_t = np.array(a[0] + a[1] * x_ref + a[2] * y_ref)
    m_w = np.array(b[3] + b[4] * x_ref + b[5] * y_ref + b[6] * xx + b[7] * x_ref * y_ref + b[8] * yy)
    c_w = np.array(b[0] + b[1] * x_ref) + b[2] * y_ref
    return m_t, c_t, m_w, c_w


class _SpectrumTrace(object):
    """Straight-line trace and linear wavelength solution for one source
    position; valid to the right of the source (x > x_ref)."""

    def __init__(self, x_ref, y_ref, trace_coeff, wl_solution):
        self.x_ref = x_ref
        self.y_ref = y_ref
        self.trace_coeff = trace_coeff
        self.wl_solution = wl_solution
        self.m_t, self.c_t, self.m_w, self.c_w = \
            self._get_wavelength_calibration_coeffs(x_ref, y_ref)
        self.m_wl, self.c_wl = self._get_x_to_wl_poly_coeffs(x_ref, y_ref)

    def _get_wavelength_calibration_coeffs(self, x_ref, y_ref):
        return wavelength_calibration_coeffs(x_ref, y_ref, self.trace_coeff, self.wl_solution)

    def x_to_y(self, x):
        return self.m_t * (x - self.x_ref) + self.c_t + self.y_ref

    def y_to_x(self, y):
        return ((y - self.y_ref - self.c_t) / self.m_t) + self.x_ref

    def _get_x_to_wl_poly_coeffs(self, x_ref, y_ref):
        """lambda(x) = m_wl x + c_wl [micron] through two probe points 10 and 20
        pixels right of the source, measured along the trace."""
        x = np.array([x_ref + 10, x_ref + 20])
        y = self.x_to_y(x)
        d = np.sqrt((y - y_ref) ** 2 + (x - x_ref) ** 2)
        wl = (self.m_w * d + self.c_w) * ANGSTROM_TO_MICRON
        m_wl = (wl[1] - wl[0]) / (x[1] - x[0])
        c_wl = wl[0] - m_wl * x[0]
        return m_wl, c_wl

    def x_to_wl(self, x):
        return self.m_wl * x + self.c_wl

    def y_to_wl(self, y):
        return self.x_to_wl(self.y_to_x(y))

    def wl_to_x(self, wl):
        return (u.value_in(wl, u.micron) - self.c_wl) / self.m_wl

    def wl_to_y(self, wl):
        return self.x_to_y(self.wl_to_x(wl))

    def psf_line(self, wl):
        x = self.wl_to_x(wl)
        y = self.wl_to_y(wl)
        m = -np.array(1.) / self.m_t
        return x, y, m, y - m * x

    def xangle(self):
        x = np.array([1., 2.])
        y = self.x_to_y(x)
        return np.arctan((y[1] - y[0]) / (x[1] - x[0]))

    def psf_length_per_pixel(self):
        return 1 / np.cos(self.xangle())


class G141_Trace(_SpectrumTrace):
    def __init__(self, x_ref, y_ref):
        _SpectrumTrace.__init__(self, x_ref, y_ref, g141_trace_coeff, g141_wl_solution)
        self.grism_name = 'G141'


class G102_Trace(_SpectrumTrace):
    def __init__(self, x_ref, y_ref):
        _SpectrumTrace.__init__(self, x_ref, y_ref, g102_trace_coeff, g102_wl_solution)
        self.grism_name = 'G102'


class G141(object):
    """WFC3 G141 grism (also the base of G102)."""

    FLAT_FILE = 'WFC3.IR.G141.flat.2.fits'
    SKY_FILE = 'WFC3.IR.G141.sky.V1.0.fits'
    SENS_FILE = 'WFC3.IR.G141.1st.sens.2.fits'

    def __init__(self):
        self.detector = WFC3_IR()
        self.name = 'G141'
        self.trace = G141_Trace

        self.min_lambda = 1.075 * u.micron
        self.max_lambda = 1.7 * u.micron
        self.resolution = 130

        self.trace_coeff = g141_trace_coeff
        self.wl_solution = g141_wl_solution

        self.flat_file_name = G141.FLAT_FILE
        self.sky_file_name = self.SKY_FILE
        self.throughput_file_name = self.SENS_FILE
        self._flat = None
        self._sky = None
        self._sens = None

        self.psf_ratio_poly = np.poly1d(PSF_RATIO_POLY)
        self.psf_sigmal_poly = np.poly1d(PSF_SIGMAL_POLY)
        self.psf_sigmah_poly = np.poly1d(PSF_SIGMAH_POLY)

        # just outside the real band so the PSF wings are not cropped
        self.wl_limits = (0.988 * u.micron, 1.777 * u.micron)

        self._FWHM_to_StDev = 1. / (2 * np.sqrt(2 * np.log(2)))

    # -- calibration files (lazy, cached) -----------------------------------
    @property
    def flat_file(self):
        return params.calb_path(self.flat_file_name)

    @property
    def sky_file(self):
        return params.calb_path(self.sky_file_name)

    @property
    def throughput_file(self):
        return params.calb_path(self.throughput_file_name)

    def _load_flat(self):
        if self._flat is None:
            # NOTE: G102 inherits this and therefore the G141 cube,
            # exactly like the reference; see G102(use_own_flat=...)
            with fits.open(params.calb_path(self.flat_file_name)) as f:
                self._flat = {
                    'wmin': f[0].header['WMIN'], 'wmax': f[0].header['WMAX'],
                    'f': tuple(f[i].data for i in range(4)),
                }
        return self._flat

    flat_wmin = property(lambda self: self._load_flat()['wmin'])
    flat_wmax = property(lambda self: self._load_flat()['wmax'])
    flat_f0 = property(lambda self: self._load_flat()['f'][0])
    flat_f1 = property(lambda self: self._load_flat()['f'][1])
    flat_f2 = property(lambda self: self._load_flat()['f'][2])
    flat_f3 = property(lambda self: self._load_flat()['f'][3])

    def _load_sens(self):
        if self._sens is None:
            with fits.open(params.calb_path(self.throughput_file_name)) as f:
                tbl = f[1].data
                wl = tbl['WAVELENGTH'] * ANGSTROM_TO_MICRON     # float32 * float -> float64
                self._sens = (np.asarray(wl, dtype=np.float64),
                              np.asarray(tbl['SENSITIVITY'], dtype=np.float64))
        return self._sens

    @property
    def throughput_wl(self):
        return self._load_sens()[0] * u.micron

    @property
    def throughput_val(self):
        return self._load_sens()[1]

    # -- wavelength-only tables ----------------------------------------------
    def set_current_wavelength_only_dependent_array(self, wl):
        wl = np.asarray(u.value_in(wl, u.micron), dtype=float)
        sens_wl, sens_val = self._load_sens()
        self.current_psf_ratio = self.psf_ratio_poly(wl)
        self.current_psf_sigmal = self.psf_sigmal_poly(wl)
        self.current_psf_sigmah = self.psf_sigmah_poly(wl)
        self.current_throughput_interpolated_function = np.interp(wl, sens_wl, sens_val)

    def apply_throughput(self, wl, flux):
        sens_wl, sens_val = self._load_sens()
        return flux * np.interp(np.asarray(u.value_in(wl, u.micron), dtype=float), sens_wl, sens_val)

    # -- trace ------------------------------------------------------------------
    def _get_wavelength_calibration_coeffs(self, x_ref, y_ref):
        return wavelength_calibration_coeffs(x_ref, y_ref, self.trace_coeff, self.wl_solution)

    def get_trace(self, x_ref, y_ref):
        return self.trace(x_ref, y_ref)

    def _pixel_wl(self, x_ref, y_ref, x, y):
        """Wavelength [angstrom] of detector position (x, y): dispersion applied
        to the perpendicular distance from the line through the source normal to
        the trace."""
        m_t, _, m_w, c_w = self._get_wavelength_calibration_coeffs(x_ref, y_ref)
        inv = 1 / m_t
        d = np.sqrt((y_ref - y + inv * x_ref - inv * x) ** 2 / (inv ** 2 + 1))
        return m_w * d + c_w

    def get_pixel_wl(self, x_ref, y_ref, x_1, y_1):
        return self._pixel_wl(x_ref, y_ref, x_1, y_1)

    def get_pixel_wl_per_row(self, x_ref, y_ref, x_values=None, y_value=None):
        x_values = np.arange(1014) if x_values is None else np.array(x_values)
        if y_value is None:
            y_value = y_ref
        return self._pixel_wl(x_ref, y_ref, x_values, y_value)

    def get_pixel_wl_whole_detector(self, x_ref, y_ref):
        ys, xs = np.mgrid[0:1014, 0:1014]
        return self._pixel_wl(x_ref, y_ref, xs, ys)

    def _bin_centers_to_limits(self, centers, bin_size=1.):
        centers = np.array(centers)
        half = bin_size / 2.
        return np.append(centers - half, centers[-1] + half)

    def get_pixel_edges_wl_per_row(self, x_ref, y_ref, x_centers=None, y_value=None,
                                   pixel_size=1.):
        return self.get_pixel_wl_per_row(
            x_ref, y_ref, self._bin_centers_to_limits(x_centers, pixel_size), y_value)

    # -- flat / sky -----------------------------------------------------------
    def flat_offset(self, size):
        """Index offset between a SUBARRAY-sized array and the 1014 flat planes
        (py2 floor division, grism.py:361-363; -5 for the 1024 full frame)."""
        return (1014 - size) // 2 if size is not None else 0

    def get_flat_field(self, x_ref, y_ref, size=None, indices=None):
        """Wavelength-dependent flat for a source at (x_ref, y_ref).

        With ``indices`` (a ``np.where`` tuple in the cropped array's
        coordinates) only those pixels are evaluated and the rest is 1.
        Host/numpy version; the exposure path evaluates the same expression in
        csrc/gather.cuh."""
        fl = self._load_flat()
        f0, f1, f2, f3 = fl['f']
        n = len(f0)
        if indices is not None:
            off = self.flat_offset(size)
            ys = np.asarray(indices[0]) + off
            xs = np.asarray(indices[1]) + off
            ys = np.where(ys < 0, ys + n, ys)     # numpy negative-index wrap (1024 frame)
            xs = np.where(xs < 0, xs + n, xs)
            sel = (ys, xs)
        else:
            ys, xs = np.mgrid[0:n, 0:n]
            sel = (slice(None), slice(None))
        w = self._pixel_wl(x_ref, y_ref, xs, ys)
        t = (w - fl['wmin']) / (fl['wmax'] - fl['wmin'])
        t2 = t * t
        t3 = t2 * t
        value = f0[sel] + (f1[sel] * t) + (f2[sel] * t2) + (f3[sel] * t3)
        if indices is None:
            flat = value
            return tools.crop_central_box(flat, size) if size is not None else flat
        # the flat value of hit pixel (r, c) is applied to pixel (r, c) itself
        side = size if size is not None and size < n else n
        flat = np.ones((side, side))
        flat[(np.asarray(indices[0]), np.asarray(indices[1]))] = value
        return flat

    def get_master_sky(self, size=None):
        if self._sky is None:
            with fits.open(params.calb_path(self.sky_file_name)) as f:
                self._sky = f[0].data
        sky = self._sky.copy()
        if size is not None:
            sky = tools.crop_central_box(sky, size)
        return sky


class G102(G141):
    SKY_FILE = 'WFC3.IR.G102.sky.V1.0.fits'
    SENS_FILE = 'WFC3.IR.G102.1st.sens.2.fits'

    def __init__(self, use_own_flat=False):
        """``use_own_flat=False`` is faithful to the reference, which loads the
        G141 flat cube for G102 and only rebinds the file name (grism.py:426-455,
        SURVEY B3); ``True`` loads WFC3.IR.G102.flat.2.fits instead."""
        G141.__init__(self)
        self.name = 'G102'
        self.trace = G102_Trace
        self.min_lambda = 0.8 * u.micron
        self.max_lambda = 1.15 * u.micron
        self.resolution = 210
        self.trace_coeff = g102_trace_coeff
        self.wl_solution = g102_wl_solution
        if use_own_flat:
            self.flat_file_name = 'WFC3.IR.G102.flat.2.fits'
        self.wl_limits = (0.75 * u.micron, 1.2 * u.micron)
