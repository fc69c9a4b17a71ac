"""WFC3 IR detector model: mode tables, read times, calibration planes.

Host-side mirror of ``wayne.detector.WFC3_IR`` (wayne/detector.py:16-350) with
the same public names and error behaviour.  The per-pixel arithmetic
(dark current, non-linearity, read noise, gain) runs in the fused CUDA pass
(csrc/reads.cuh); the numpy versions kept here (`add_dark_current`,
`apply_non_linearity`, `add_read_noise`) serve callers that use the class
directly and are NOT used by ExposureGenerator.

Differences from the reference, all deliberate:
 * the mode tables are read from ``data/wfc3_ir_modes.json`` (converted from the
   reference's CSVs by tools/import_reference_data.py) -- no pandas needed;
 * calibration FITS files are opened lazily, once, and cached (the reference
   re-opens the gain file on every read, detector.py:200-209);
 * ``*_planes`` helpers prepare float64 device planes using the SAME numpy
   dtype arithmetic the reference's expressions perform on the FITS arrays
   (float32 planes stay float32 until the reference would promote them).
"""
from __future__ import annotations

import json
import os

import numpy as np

from . import fitsio as fits
from . import params, tools
from . import units as u


class WFC3SimException(BaseException):
    pass


class WFC3SimSampleModeError(WFC3SimException):
    pass


class WFC3SimNoDarkFileError(WFC3SimException):
    pass


class WFC3_IR(object):
    """Methods and calibrations of the WFC3 IR channel."""

    def __init__(self):
        self._pixel_array = np.zeros((1024, 1024))
        self.pixel_size_micron = 18.0
        self.telescope_area = np.pi * (2.4 / 2.) ** 2 * (u.m ** 2)
        self.min_counts = -20
        self.max_counts = 78000  # DN, 5% non-linearity limit

        self.constant_gain = 2.35
        self.gain_file_name = 'u4m1335mi_pfl.fits'
        self.read_noise = 14.1 / self.constant_gain  # e- -> DN

        self.initial_bias = os.path.join(params._data_dir, 'wfc3_ir_initial_bias_256.npz')

        self.telescope = 'HST'
        self.instrument = 'WFC3'
        self.detector_type = 'IR'

        self.modes_exp_table, self.modes_calb_table = self._get_modes()

        self.non_linear_file_name = 'u1k1727mi_lin.fits'
        self._nl = None
        self._gain_raw = None
        self._dark_cache = {}

    # ------------------------------------------------------------------
    # calibration files (lazy)
    # ------------------------------------------------------------------
    @property
    def gain_file(self):
        return os.path.join(params._calb_dir, self.gain_file_name)

    @property
    def non_linear_file(self):
        return os.path.join(params._calb_dir, self.non_linear_file_name)

    def _load_nl(self):
        if self._nl is None:
            with fits.open(params.calb_path(self.non_linear_file_name)) as f:
                self._nl = tuple(f[i].data for i in (1, 2, 3, 4))
        return self._nl

    non_linear_c1 = property(lambda self: self._load_nl()[0])
    non_linear_c2 = property(lambda self: self._load_nl()[1])
    non_linear_c3 = property(lambda self: self._load_nl()[2])
    non_linear_c4 = property(lambda self: self._load_nl()[3])

    def get_initial_bias(self):
        """266 x 266 float64 initial bias used as the zero read of SUBARRAY=256
        exposures (exposure_generator.py:456-458)."""
        if getattr(self, '_bias', None) is None:
            self._bias = np.load(self.initial_bias)['bias'].astype(np.float64)
        return self._bias.copy()

    # ------------------------------------------------------------------
    # mode tables
    # ------------------------------------------------------------------
    def _mode_rows(self):
        if not hasattr(self, '_rows'):
            with open(os.path.join(params._data_dir, 'wfc3_ir_modes.json')) as f:
                doc = json.load(f)
            self._rows = doc
            self._exp_index = {}
            for sub, seq, num, t in doc['exptime']:
                self._exp_index.setdefault((sub, seq), []).append((num, t))
            self._dark_index = {(sub, seq): name for sub, seq, name in doc['dark']}
        return self._rows

    def _get_modes(self):
        """(exposure-time table, calibration-file table); pandas DataFrames when
        pandas is importable (shapes (360, 4) and (19, 3)), else lists of rows."""
        doc = self._mode_rows()
        try:
            import pandas as pd
        except ImportError:
            return doc['exptime'], doc['dark']
        return (pd.DataFrame(doc['exptime'], columns=doc['exptime_columns']),
                pd.DataFrame(doc['dark'], columns=doc['dark_columns']))

    def exptime(self, NSAMP, SUBARRAY, SAMPSEQ):
        """Total exposure time of a mode (tables quote SAMPNUM = NSAMP - 1)."""
        self._mode_rows()
        sample_number = NSAMP - 1
        for num, t in self._exp_index.get((SUBARRAY, SAMPSEQ), ()):
            if num == sample_number:
                return t * u.s
        raise WFC3SimSampleModeError(
            "SAMPSEQ = {}, NSAMP={}, SUBARRAY={} is not a permitted combination"
            "".format(SAMPSEQ, NSAMP, SUBARRAY))

    def get_read_times(self, NSAMP, SUBARRAY, SAMPSEQ):
        """Time of every non-zero read up to NSAMP, in table order."""
        if not 2 <= NSAMP <= 16:
            raise WFC3SimSampleModeError(
                "NSAMP must be an integer between 2 and 16, got {}".format(NSAMP))
        self._mode_rows()
        sample_number = NSAMP - 1
        times = [t for num, t in self._exp_index.get((SUBARRAY, SAMPSEQ), ())
                 if num <= sample_number]
        if not times:
            raise WFC3SimSampleModeError(
                "SAMPSEQ = {}, NSAMP={}, SUBARRAY={}  is not a permitted "
                "combination".format(SAMPSEQ, NSAMP, SUBARRAY))
        return np.array(times) * u.s

    def num_exp_per_buffer(self, NSAMP, SUBARRAY):
        """Exposures that fit before a buffer dump (304 headers / 2 full-frame
        16-read exposures; py2 integer division as in detector.py:287)."""
        hard_limit = 304
        headers_per_exp = NSAMP + 1
        total_allowed_reads = 2 * 16 * (1024 // SUBARRAY)
        if total_allowed_reads > hard_limit:
            total_allowed_reads = hard_limit
        return int(np.floor(total_allowed_reads / headers_per_exp))

    # ------------------------------------------------------------------
    # geometry
    # ------------------------------------------------------------------
    @staticmethod
    def light_side(subarray):
        return 1014 if subarray == 1024 else subarray

    @staticmethod
    def full_side(subarray):
        return min(subarray + 10, 1024)

    def gen_pixel_array(self, subarray, light_sensitive=True):
        n = self.light_side(subarray) if light_sensitive else self.full_side(subarray)
        return np.zeros((n, n))

    def add_bias_pixels(self, pixel_array):
        allowed_input = (1014, 512, 256, 128, 64)
        n = len(pixel_array)
        if n not in allowed_input:
            raise ValueError('array size must be in {} got {}'.format(allowed_input, n))
        full = np.zeros((n + 10, n + 10))
        full[5:-5, 5:-5] = pixel_array
        return full

    # ------------------------------------------------------------------
    # dark current
    # ------------------------------------------------------------------
    def _dark_file(self, SUBARRAY, SAMPSEQ):
        self._mode_rows()
        try:
            return self._dark_index[(SUBARRAY, SAMPSEQ)]
        except KeyError:
            raise WFC3SimNoDarkFileError(
                "No Dark file found for SAMPSEQ = {}, SUBARRAY={}".format(SAMPSEQ, SUBARRAY))

    def dark_planes(self, NSAMP, SUBARRAY, SAMPSEQ):
        """(dark, dark_error) of the read with that NSAMP: extensions
        ``-(NSAMP)*5`` and ``+1`` of the mode's super-dark (detector.py:183-190);
        the error plane already has its non-positive entries replaced by 1e-5."""
        name = self._dark_file(SUBARRAY, SAMPSEQ)
        key = (name, NSAMP)
        if key not in self._dark_cache:
            try:
                path = params.calb_path(name)
            except params.CalibrationFileMissing as exc:
                raise WFC3SimNoDarkFileError(str(exc))
            if name not in self._dark_cache:
                self._dark_cache[name] = fits.open(path)
            f = self._dark_cache[name]
            idx = -(NSAMP) * 5
            dark = f[idx].data
            err = f[idx + 1].data
            self._dark_cache[key] = (dark, np.where(err > 0, err, 0.00001))
        return self._dark_cache[key]

    def add_dark_current(self, pixel_array, NSAMP, SUBARRAY, SAMPSEQ):
        dark, err = self.dark_planes(NSAMP, SUBARRAY, SAMPSEQ)
        return pixel_array + np.random.normal(dark, err)

    def add_read_noise(self, pixel_array):
        return np.random.normal(pixel_array, self.read_noise)

    # ------------------------------------------------------------------
    # gain
    # ------------------------------------------------------------------
    def get_gain(self, size):
        """2.35 / pixel-flat, light-sensitive area, cropped centrally to ``size``."""
        if self._gain_raw is None:
            with fits.open(params.calb_path(self.gain_file_name)) as f:
                self._gain_raw = f[1].data[5:-5, 5:-5]
        gain = self.constant_gain / self._gain_raw
        if size is not None:
            gain = tools.crop_central_box(gain, size)
        return gain

    # ------------------------------------------------------------------
    # non-linearity
    # ------------------------------------------------------------------
    def _nl_cropped(self, n):
        c = self._load_nl()
        half = len(c[0]) // 2
        lo, hi = half - n // 2, half + n // 2
        return tuple(p[lo:hi, lo:hi] for p in c)

    def non_linear_planes(self, n):
        """The seven float64 planes the CUDA Newton solve reads: 1+c1, c2, c3,
        c4, 2*c2, 3*c3, 4*c4 -- formed in the calibration file's own dtype, as
        numpy does inside detector.py:339-343."""
        c1, c2, c3, c4 = self._nl_cropped(n)
        planes = (1 + c1, c2, c3, c4, 2 * c2, 3 * c3, 4 * c4)
        return tuple(np.ascontiguousarray(p, dtype=np.float64) for p in planes)

    def apply_non_linearity(self, pixel_array):
        """Inverse non-linearity by Newton-Raphson with the reference's global
        stopping rule (all pixels move < 1e-3, at most 10^4 steps)."""
        c1, c2, c3, c4 = self._nl_cropped(len(pixel_array))
        u0 = pixel_array
        u1 = u0 * 0
        for _ in range(10000):
            u1 = u0 - ((-pixel_array + u0 * (1 + c1 + u0 * (c2 + u0 * (c3 + c4 * u0)))) /
                       (1 + c1 + 2 * c2 * u0 + 3 * c3 * u0 * u0 + 4 * c4 * u0 * u0 * u0))
            if (np.abs(u1 - u0) < 10 ** (-3)).all():
                break
            u0 = u1
        return u1

    def apply_quantum_efficiency(self, wl, counts):
        raise NotImplementedError(
            "quantum efficiency is folded into the grism sensitivity curve and is not on the "
            "exposure path (wayne/exposure_generator.py:614)")
