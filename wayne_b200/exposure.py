"""Exposure: the result container of the hot path and its FITS writer.

Mirror of ``wayne.exposure.Exposure`` (wayne/exposure.py:22-429): ``reads`` is
a list of ``(ndarray, header)`` with the zero read first, ``add_read`` appends,
``generate_fits`` writes an HST-style multi-extension file (five HDUs per read,
reads in reverse order, SAMPNUM / SAMPTIME / DELTATIM / CRPIX1 per read).

The per-read post-processing methods of the reference (non-linearity, dark,
clip, reference-pixel reset, zero read, read noise -- exposure.py:49-131) are
fused into the CUDA per-pixel pass (csrc/reads.cuh); numpy versions with the
same names are kept for callers that post-process an Exposure by hand.
"""
from __future__ import annotations

import datetime
import os
import sys

import numpy as np

from . import fitsio as fits
from . import params
from . import units as u

__version__ = "b200-0.1"


def _val(x, unit=None, default=0.0):
    if x is None or x is False:
        return default
    if u.is_quantity(x):
        return float(x.to(unit).value) if unit is not None else float(x.value)
    try:
        return float(x)
    except (TypeError, ValueError):
        return default


class Exposure(object):
    def __init__(self, detector=None, filter=None, planet=None, exp_info=None):
        self.detector = detector
        self.filter = filter
        self.planet = planet
        self.exp_info = exp_info
        self.SUBARRAY = exp_info['SUBARRAY']
        self.NSAMP = exp_info['NSAMP']
        self.SAMPSEQ = exp_info['SAMPSEQ']
        self._reads = []   # read 0 (zero read) first
        self._pending = None

    # -- container -------------------------------------------------------
    @property
    def reads(self):
        """[(ndarray, header)], zero read first.  The device path hands the
        exposure back while its device->host copy is still in flight; the first
        access waits for it (so a driver that writes the FITS file right away
        behaves exactly like the reference, and one that keeps generating
        overlaps the copy with the next exposure)."""
        if self._pending is not None:
            pending, self._pending = self._pending, None
            pending(self)
        return self._reads

    @reads.setter
    def reads(self, value):
        self._pending = None
        self._reads = value

    def add_read(self, data, read_info=None):
        header = self.generate_read_header(read_info) if read_info is not None else fits.Header()
        self._reads.append((data, header))

    # -- host versions of the per-read operations --------------------------
    def _map(self, fn, start=0):
        for i in range(start, len(self.reads)):
            arr, hdr = self.reads[i]
            self.reads[i] = (fn(i, arr), hdr)

    def apply_non_linear(self):
        self._map(lambda i, a: self.detector.apply_non_linearity(a), 1)

    def add_read_noise(self):
        self._map(lambda i, a: self.detector.add_read_noise(a))

    def add_dark_current(self):
        self._map(lambda i, a: self.detector.add_dark_current(a, i + 1, self.SUBARRAY,
                                                               self.SAMPSEQ), 1)

    def scale_counts_between_limits(self):
        lo, hi = self.detector.min_counts, self.detector.max_counts
        self._map(lambda i, a: np.clip(a, lo, hi))

    def add_zero_read(self):
        zero = self.reads[0][0]
        self._map(lambda i, a: a + zero, 1)

    def reset_reference_pixels(self, value=0.):
        def reset(i, a):
            border = np.ones_like(a, dtype=bool)
            border[5:-5, 5:-5] = False
            a[border] = value
            return a
        self._map(reset)

    # -- FITS ----------------------------------------------------------------
    def generate_read_header(self, read_info):
        h = fits.Header()
        h['CRPIX1'] = (read_info['CRPIX1'], 'x-coordinate of reference pixel')
        h['SAMPTIME'] = (_val(read_info['cumulative_exp_time'], u.s), 'total integration time (sec)')
        h['DELTATIM'] = (_val(read_info['read_exp_time'], u.s), 'sample integration time (sec)')
        return h

    def generate_science_header(self, ldcoeffs=None):
        info = self.exp_info
        h = fits.Header()
        h['DATE'] = (datetime.datetime.now().strftime("%Y-%m-%d"),
                     'date this file was written (yyyy-mm-dd)')
        h['FILENAME'] = (info['filename'], 'name of file')
        h['FILETYPE'] = ('SCI', 'type of data found in data file')
        h['TELESCOP'] = (self.detector.telescope, 'telescope used to acquire data')
        h['INSTRUME'] = (self.detector.instrument, 'identifier for instrument used to acquire data')
        h['EQUINOX'] = (2000.0, 'equinox of celestial coord. system')
        h['PRIMESI'] = (self.detector.instrument, 'instrument designated as prime')
        h['TARGNAME'] = (getattr(self.planet, 'name', 'None'), "proposer's target name")
        # exposure information (MJD like the reference: JD - 2400000.5)
        h['EXPSTART'] = (_val(info['EXPSTART'], u.day) - 2400000.5, 'exposure start time (MJD)')
        h['EXPEND'] = (_val(info['EXPEND'], u.day) - 2400000.5, 'exposure end time (MJD)')
        h['EXPTIME'] = (_val(info['EXPTIME'], u.s), 'exposure duration (seconds)')
        h['POSTARG1'] = (0., 'POSTARG in axis 1 direction')
        h['POSTARG2'] = (_val(info.get('SCAN_DIR')), 'POSTARG in axis 2 direction')
        h['OBSTYPE'] = (info['OBSTYPE'], 'observation type - imaging or spectroscopic')
        h['OBSMODE'] = ('MULTIACCUM', 'operating mode')
        h['SCLAMP'] = ('NONE', 'lamp status, NONE or name of lamp which is on')
        h['SUBARRAY'] = (bool(info['SUBARRAY'] != 1024), 'data from a subarray (T) or full frame (F)')
        h['SUBTYPE'] = ('SQ{}SUB'.format(info['SUBARRAY']), 'size/type of IR subarray')
        h['DETECTOR'] = (self.detector.detector_type, 'detector in use: UVIS or IR')
        h['FILTER'] = (getattr(self.filter, 'name', 'None'), 'element selected from filter wheel')
        h['SAMP_SEQ'] = (info['SAMPSEQ'], 'MultiAccum exposure time sequence name')
        h['NSAMP'] = (info['NSAMP'], 'number of MULTIACCUM samples')
        h['SAMPZERO'] = (0., 'sample time of the zeroth read (sec)')
        h['APERTURE'] = ('GRISM{}'.format(info['SUBARRAY']), 'aperture name')
        h['DIRIMAGE'] = ('NONE', 'direct image for grism or prism exposure')
        # simulation provenance
        h['SIM'] = (True, 'Wayne Simulation (T/F)')
        h['SIM-VER'] = (__version__, 'simulator version used')
        h['SIM-TIME'] = (_val(info.get('sim_time'), u.s), 'simulation time (s)')
        h['X-REF'] = (_val(info.get('x_ref')), 'x position of star on frame (full frame)')
        h['Y-REF'] = (_val(info.get('y_ref')), 'y position of star on frame (full frame)')
        h['SAMPRATE'] = (_val(info.get('samp_rate'), u.s), 'How often exposure is sampled (s)')
        h['NSE-MEAN'] = (_val(info.get('noise_mean')), 'mean of normal noise (per s per pix)')
        h['NSE-STD'] = (_val(info.get('noise_std')), 'std of normal noise (per s per pix)')
        h['ADD-DRK'] = (bool(info.get('add_dark')), 'dark current added (T/F)')
        h['ADD-FLAT'] = (bool(info.get('add_flat')), 'flat field added (T/F)')
        h['ADD-GAIN'] = (bool(info.get('add_gain')), 'gain variations added (T/F)')
        h['ADD-NLIN'] = (bool(info.get('add_non_linear')), 'non-linearity effects added (T/F)')
        h['STAR-NSE'] = (bool(info.get('add_stellar_noise')), 'Stellar noise added (T/F)')
        h['CSMCRATE'] = (_val(info.get('cosmic_rate')), 'Rate of cosmic hits (per s)')
        h['SKY-LVL'] = (_val(info.get('sky_background'), u.count / u.s), 'multiple of master sky per s')
        h['VSTTREND'] = (_val(info.get('scale_factor'), default=1.0), 'visit trend scale factor')
        h['CLIPVALS'] = (bool(info.get('clip_values_det_limits')), 'pixels clipped to detector range (T/F)')
        h['RANDSEED'] = (params.seed if params.seed is not None else -1, 'seed used for the visit')
        h['RNG'] = (str(info.get('rng', 'philox')), 'random stream: philox (native) or numpy (compat)')
        h['V-PY'] = ('.'.join(str(v) for v in sys.version_info[:3]), 'Python version used')
        h['V-NP'] = (np.__version__, 'NumPy version used')
        if ldcoeffs is not None:
            for i, c in enumerate(ldcoeffs, 1):
                h['LD%d' % i] = (float(c), 'Non-linear limb darkening coeff %d' % i)
        h['STARX'] = (_val(info.get('x_ref')), 'x position of star on frame (full frame))')
        return h

    def generate_fits(self, out_dir='', filename=None, ldcoeffs=None):
        """HST-style file: primary header, then per read (last read first)
        SCI, ERR, DQ, SAMP, TIME extensions (the latter four empty, as in the
        reference)."""
        assert len(self.reads) == self.exp_info['NSAMP'], \
            'Reads {} != NSAMP {}'.format(len(self.reads), self.exp_info['NSAMP'])
        if filename is None:
            filename = self.exp_info['filename']
        out_path = os.path.join(out_dir, filename)
        hdus = [fits.HDU(None, self.generate_science_header(ldcoeffs=ldcoeffs))]
        n = len(self.reads)
        for i, (data, header) in enumerate(reversed(self.reads)):
            hdr = fits.Header(header)
            hdr.comments = dict(getattr(header, 'comments', {}))
            hdr['SAMPNUM'] = n - 1 - i
            hdr['EXTNAME'] = 'SCI'
            hdr['EXTVER'] = i + 1
            hdus.append(fits.HDU(np.asarray(data), hdr))
            for name in ('ERR', 'DQ', 'SAMP', 'TIME'):
                eh = fits.Header()
                eh['EXTNAME'] = name
                eh['EXTVER'] = i + 1
                hdus.append(fits.HDU(None, eh))
        fits.writeto(out_path, hdus, overwrite=True)
        return out_path
