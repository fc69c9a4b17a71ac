"""Shared test inputs: seeded synthetic spectra and the calibration planes as
plain arrays (for the oracle) read from the same synthetic files the product
reads.  Test infrastructure."""
import numpy as np

from wayne_b200 import fitsio, params


def oracle_calibration(grism_name='G141', dark_mode=(256, 'SPARS10'), nsamp=5):
    """dict of arrays in the dtypes the FITS files hold (oracle/exposure_oracle.py)."""
    from wayne_b200.detector import WFC3_IR
    det = WFC3_IR()
    cal = {}
    # the reference loads the G141 flat cube for both grisms
    with fitsio.open(params.calb_path('WFC3.IR.G141.flat.2.fits')) as f:
        cal['flat'] = tuple(f[i].data for i in range(4))
        cal['flat_wmin'], cal['flat_wmax'] = f[0].header['WMIN'], f[0].header['WMAX']
    with fitsio.open(params.calb_path('WFC3.IR.%s.sky.V1.0.fits' % grism_name)) as f:
        cal['sky'] = f[0].data
    with fitsio.open(params.calb_path('WFC3.IR.%s.1st.sens.2.fits' % grism_name)) as f:
        t = f[1].data
        cal['sens_wl_um'] = np.asarray(t['WAVELENGTH'] * 1e-4, dtype=np.float64)
        cal['sens_val'] = np.asarray(t['SENSITIVITY'], dtype=np.float64)
    with fitsio.open(params.calb_path('u4m1335mi_pfl.fits')) as f:
        cal['pfl'] = f[1].data
    with fitsio.open(params.calb_path('u1k1727mi_lin.fits')) as f:
        cal['nl'] = tuple(f[i].data for i in (1, 2, 3, 4))
    cal['bias256'] = det.get_initial_bias()
    if dark_mode is not None:
        try:
            name = det._dark_file(*dark_mode)
            with fitsio.open(params.calb_path(name)) as f:
                cal['dark'] = {n: (f[-n * 5].data, f[-n * 5 + 1].data) for n in range(2, nsamp + 1)}
        except Exception:
            cal['dark'] = None
    return cal


def spectrum(n_wl=600, lo=0.9, hi=1.9, level=1.0e-13, seed=11, depth=0.0146):
    """wl [micron], stellar flux [erg/(A s cm^2)], per-bin transit depth."""
    rng = np.random.default_rng(seed)
    wl = np.linspace(lo, hi, n_wl)
    x = 1.4388e4 / (wl * 6065.0)
    bb = 1.0 / (wl ** 5 * (np.exp(x) - 1.0))
    flux = level * bb / bb.max() * (1 + 0.02 * rng.standard_normal(n_wl))
    planet = depth * (1 + 0.01 * np.sin(12 * wl) + 0.002 * rng.standard_normal(n_wl))
    return wl, flux, planet
