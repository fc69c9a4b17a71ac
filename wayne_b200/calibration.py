"""Synthetic, seeded stand-ins for the WFC3 calibration set.

The real calibration files are downloaded by the reference at import time
(wayne/params.py:26-56) and are not redistributable here (no network in the
build or on the GPU box).  This module writes files with the SAME names, HDU
layout and dtypes the reference's readers expect (wayne/grism.py:66-106,
453-476; wayne/detector.py:31-67, 172-191, 200-209), filled with seeded
synthetic planes, so that the CPU oracle and the CUDA path read
identical inputs through identical code paths:

  WFC3.IR.G141.flat.2.fits / WFC3.IR.G102.flat.2.fits   4 x 1014^2 float32 cube, WMIN/WMAX
  WFC3.IR.G141.sky.V1.0.fits / ...G102...               1014^2 float32
  WFC3.IR.G141.1st.sens.2.fits / ...G102...             BINTABLE WAVELENGTH[A], SENSITIVITY
  u4m1335mi_pfl.fits                                    ext 1: 1024^2 float32 pixel flat
  u1k1727mi_lin.fits                                    ext 1-4: 1024^2 float32 c1..c4
  <mode>_drk.fits                                       16 reads x (SCI, ERR, DQ, SAMP, TIME)
"""
from __future__ import annotations

import os

import numpy as np

from . import fitsio as fits
from .detector import WFC3_IR

DEFAULT_SEED = 20170410


def _hdu(data=None, name=None, **cards):
    h = fits.HDU(data, fits.Header(), name)
    for k, v in cards.items():
        h.header[k] = v
    return h


def _sens_curve(wl_a, lo_um, hi_um, peak):
    um = wl_a * 1e-4
    edge = 0.012
    rise = 1.0 / (1.0 + np.exp(-(um - lo_um) / edge))
    fall = 1.0 / (1.0 + np.exp((um - hi_um) / edge))
    tilt = 0.75 + 0.25 * (um - lo_um) / (hi_um - lo_um)
    return peak * rise * fall * tilt


def write_synthetic_calibration(dirpath, modes=((256, 'SPARS10'),), seed=DEFAULT_SEED,
                                overwrite=False):
    """Write the calibration set into ``dirpath``; returns ``dirpath``.

    ``modes``: (SUBARRAY, SAMPSEQ) pairs to write super-darks for (a 1024 dark
    is 128 MB, so only what is asked for is written).  Deterministic in ``seed``.
    """
    os.makedirs(dirpath, exist_ok=True)
    rng = np.random.default_rng(seed)
    f32 = np.float32

    def path(name):
        return os.path.join(dirpath, name)

    def need(name):
        return overwrite or not os.path.isfile(path(name))

    # draw everything in a fixed order so the set does not depend on what exists
    flat = [(1 + 0.01 * rng.standard_normal((1014, 1014))).astype(f32)]
    flat += [(0.005 * rng.standard_normal((1014, 1014))).astype(f32) for _ in range(3)]
    sky141 = (1 + 0.05 * rng.standard_normal((1014, 1014))).astype(f32)
    sky102 = (1 + 0.05 * rng.standard_normal((1014, 1014))).astype(f32)
    pfl = (1 + 0.02 * rng.standard_normal((1024, 1024))).astype(f32)
    c2 = (6.4e-7 * (1 + 0.05 * rng.standard_normal((1024, 1024)))).astype(f32)
    flat102 = [(1 + 0.01 * rng.standard_normal((1014, 1014))).astype(f32)]
    flat102 += [(0.005 * rng.standard_normal((1014, 1014))).astype(f32) for _ in range(3)]

    for name, cube, wmin, wmax in (('WFC3.IR.G141.flat.2.fits', flat, 9880.0, 17770.0),
                                   ('WFC3.IR.G102.flat.2.fits', flat102, 7500.0, 12000.0)):
        if need(name):
            hdus = [_hdu(cube[0], None, WMIN=wmin, WMAX=wmax)]
            hdus += [_hdu(c, 'F%d' % i) for i, c in enumerate(cube[1:], 1)]
            fits.writeto(path(name), hdus)
    for name, plane in (('WFC3.IR.G141.sky.V1.0.fits', sky141), ('WFC3.IR.G102.sky.V1.0.fits', sky102)):
        if need(name):
            fits.writeto(path(name), [_hdu(plane)])
    for name, lo, hi, a0, a1, peak in (('WFC3.IR.G141.1st.sens.2.fits', 1.08, 1.69, 10000., 18000., 4.5e16),
                                       ('WFC3.IR.G102.1st.sens.2.fits', 0.80, 1.15, 7000., 12500., 2.5e16)):
        if need(name):
            wl = np.arange(a0, a1 + 1, 10.0)
            sens = _sens_curve(wl, lo, hi, peak)
            tbl = fits.table_hdu({'WAVELENGTH': wl.astype(f32), 'SENSITIVITY': sens.astype(f32),
                                  'ERROR': (0.01 * sens).astype(f32)}, name='SENS')
            fits.writeto(path(name), [_hdu(), tbl])
    if need('u4m1335mi_pfl.fits'):
        fits.writeto(path('u4m1335mi_pfl.fits'), [_hdu(), _hdu(pfl, 'SCI')])
    if need('u1k1727mi_lin.fits'):
        zero = np.zeros((1024, 1024), f32)
        fits.writeto(path('u1k1727mi_lin.fits'),
                     [_hdu(), _hdu(zero, 'COEF1'), _hdu(c2, 'COEF2'), _hdu(zero, 'COEF3'),
                      _hdu(zero, 'COEF4')])

    det = WFC3_IR()
    for sub, seq in modes:
        name = det._dark_file(sub, seq)
        if not need(name):
            continue
        side = det.full_side(sub)
        times = {0: 0.0}
        for num, t in det._exp_index.get((sub, seq), ()):
            times[num] = t
        hdus = [_hdu()]
        for k in range(15, -1, -1):            # last read first; the zero read closes the file
            t = times.get(k, times[max(times)])
            sci = np.full((side, side), 0.05 * t, f32)
            err = np.full((side, side), 0.5, f32)
            hdus += [_hdu(sci, 'SCI', EXTVER=k + 1, SAMPNUM=k, SAMPTIME=float(t)),
                     _hdu(err, 'ERR', EXTVER=k + 1), _hdu(None, 'DQ', EXTVER=k + 1),
                     _hdu(None, 'SAMP', EXTVER=k + 1), _hdu(None, 'TIME', EXTVER=k + 1)]
        fits.writeto(path(name), hdus)
    return dirpath
