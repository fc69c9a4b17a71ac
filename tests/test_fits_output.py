"""FITS output of an exposure: Exposure.generate_fits writes
the HST-style layout of the reference (wayne/exposure.py:133-214 -- primary
science header, then per read, LAST read first: SCI, ERR, DQ, SAMP, TIME) with
this package's own FITS writer, and the reader round-trips it."""
import numpy as np

from wayne import detector, exposure, fitsio, grism
from wayne import units as u


def _exposure(nsamp=4, side=266):
    info = {'filename': '0007_raw.fits', 'EXPSTART': 2456196.3 * u.day, 'EXPEND': 2456196.301 * u.day,
            'EXPTIME': 22.3 * u.s, 'SCAN': True, 'SCAN_DIR': 1, 'OBSTYPE': 'SPECTROSCOPIC', 'NSAMP': nsamp,
            'SAMPSEQ': 'SPARS10', 'SUBARRAY': 256, 'samp_rate': 0.01 * u.s, 'sim_time': 1.5 * u.s,
            'x_ref': 404.5, 'y_ref': 457.4, 'add_dark': True, 'add_flat': True, 'add_gain': True,
            'add_non_linear': True, 'add_stellar_noise': True, 'cosmic_rate': 11.,
            'sky_background': 5.5 * u.count / u.s, 'scale_factor': 0.9991, 'clip_values_det_limits': True,
            'noise_mean': False, 'noise_std': False}
    exp = exposure.Exposure(detector.WFC3_IR(), grism.G141(), None, info)
    rng = np.random.default_rng(1)
    t = 0.0
    for r in range(nsamp):
        exp.add_read(rng.normal(100. * r, 5., (side, side)),
                     {'cumulative_exp_time': t * u.s, 'read_exp_time': (7.3 if r else 0.0) * u.s, 'CRPIX1': 0})
        t += 7.3
    return exp


def test_generate_fits_layout_and_roundtrip(tmp_path):
    exp = _exposure()
    path = exp.generate_fits(str(tmp_path), ldcoeffs=(0.1, 0.2, 0.3, 0.4))
    assert path.endswith('0007_raw.fits')
    f = fitsio.open(path)
    assert len(f) == 1 + 5 * 4
    h = f[0].header
    assert h['TELESCOP'] == 'HST' and h['INSTRUME'] == 'WFC3' and h['DETECTOR'] == 'IR'
    assert h['FILTER'] == 'G141' and h['NSAMP'] == 4 and h['SAMP_SEQ'] == 'SPARS10'
    assert h['SUBARRAY'] is True and h['SUBTYPE'] == 'SQ256SUB' and h['SIM'] is True
    assert abs(h['EXPSTART'] - (2456196.3 - 2400000.5)) < 1e-9 and abs(h['EXPTIME'] - 22.3) < 1e-12
    assert abs(h['LD3'] - 0.3) < 1e-12 and abs(h['CSMCRATE'] - 11.) < 1e-12 and abs(h['SKY-LVL'] - 5.5) < 1e-12
    names = [hdu.header.get('EXTNAME') for hdu in f[1:]]
    assert names == ['SCI', 'ERR', 'DQ', 'SAMP', 'TIME'] * 4
    # reads are stored last-first; SAMPNUM counts down to the zero read
    for i in range(4):
        sci = f[1 + 5 * i]
        assert sci.header['SAMPNUM'] == 3 - i and sci.header['EXTVER'] == i + 1
        assert np.array_equal(sci.data, exp.reads[3 - i][0])        # float64, exact
        assert abs(sci.header['SAMPTIME'] - 7.3 * (3 - i)) < 1e-9
        assert f[2 + 5 * i].data is None
    assert f[1].data.dtype.kind == 'f' and f[1].data.dtype.itemsize == 8


def test_generate_fits_checks_read_count(tmp_path):
    exp = _exposure()
    exp.reads.pop()
    try:
        exp.generate_fits(str(tmp_path))
    except AssertionError as e:
        assert 'NSAMP' in str(e)
    else:
        raise AssertionError("expected the read-count assertion of exposure.py:144-146")


def test_lazy_reads_materialise_once():
    exp = _exposure()
    calls = []

    def pending(e):
        calls.append(1)
        e.add_read(np.zeros((2, 2)))

    n = len(exp.reads)
    exp._pending = pending
    assert len(exp.reads) == n + 1 and len(exp.reads) == n + 1 and calls == [1]
