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


def test_file_against_the_fits_standard(tmp_path):
    """The file Exposure.generate_fits writes, read by tests/fits_standard.py -- a strict
    parser written from the FITS standard that shares nothing with wayne_b200.fitsio --
    and compared with the layout of the reference's writer (wayne/exposure.py:133-214,
    read headers :412-429): 2880-byte blocks, mandatory keywords in order, fixed-format
    values, big-endian BITPIX = -64 science arrays, EXTNAME sequence SCI / ERR / DQ / SAMP /
    TIME per read with the LAST read first, SAMPNUM counting down, SAMPTIME / DELTATIM /
    CRPIX1 on every SCI header."""
    import os
    import struct
    from tests import fits_standard as FS
    exp = _exposure(nsamp=5)
    path = exp.generate_fits(str(tmp_path), ldcoeffs=(0.1, 0.2, 0.3, 0.4))
    assert os.path.getsize(path) % 2880 == 0
    hdus = FS.read(path)                       # raises FitsFormatError on any deviation
    assert len(hdus) == 1 + 5 * 5
    prim, cards, data = hdus[0]
    assert data is None and prim['NAXIS'] == 0 and prim['EXTEND'] is True
    assert [k for k, _, _ in cards[:4]] == ['SIMPLE', 'BITPIX', 'NAXIS', 'EXTEND']
    # the science-header keywords the reference writes (exposure.py:216-410) that a reader keys on
    for key, want in (('TELESCOP', 'HST'), ('INSTRUME', 'WFC3'), ('DETECTOR', 'IR'), ('FILTER', 'G141'),
                      ('OBSTYPE', 'SPECTROSCOPIC'), ('OBSMODE', 'MULTIACCUM'), ('SAMP_SEQ', 'SPARS10'),
                      ('NSAMP', 5), ('SUBARRAY', True), ('SUBTYPE', 'SQ256SUB'), ('FILETYPE', 'SCI'),
                      ('FILENAME', '0007_raw.fits'), ('SIM', True)):
        assert prim[key] == want, key
    assert abs(prim['EXPSTART'] - (2456196.3 - 2400000.5)) < 1e-9 and abs(prim['EXPTIME'] - 22.3) < 1e-12
    assert [prim['LD%d' % i] for i in (1, 2, 3, 4)] == [0.1, 0.2, 0.3, 0.4]
    names = [h['EXTNAME'] for h, _, _ in hdus[1:]]
    assert names == ['SCI', 'ERR', 'DQ', 'SAMP', 'TIME'] * 5
    for i in range(5):
        h, cards, data = hdus[1 + 5 * i]
        read = 4 - i                                               # reversed(self.reads), exposure.py:160
        assert h['XTENSION'] == 'IMAGE' and h['BITPIX'] == -64 and h['NAXIS'] == 2
        assert (h['NAXIS1'], h['NAXIS2'], h['PCOUNT'], h['GCOUNT']) == (266, 266, 0, 1)
        assert h['SAMPNUM'] == read and h['EXTVER'] == i + 1 and h['CRPIX1'] == 0
        assert abs(h['SAMPTIME'] - 7.3 * read) < 1e-9 and abs(h['DELTATIM'] - (7.3 if read else 0.0)) < 1e-12
        assert data.dtype == np.dtype('>f8') and np.array_equal(data, exp.reads[read][0])
        for k in range(1, 5):                                       # the four placeholder extensions
            eh, _, ed = hdus[1 + 5 * i + k]
            assert ed is None and eh['XTENSION'] == 'IMAGE' and eh['NAXIS'] == 0 and eh['EXTVER'] == i + 1
    # byte-level spot check of the data unit: first pixel of the last read, big-endian IEEE-754
    raw = open(path, 'rb').read()
    first_sci_data = raw.index(b'XTENSION') // 2880 * 2880
    hdr_blocks = 1
    while b'END' + b' ' * 77 not in raw[first_sci_data:first_sci_data + 2880 * hdr_blocks]:
        hdr_blocks += 1
    off = first_sci_data + 2880 * hdr_blocks
    assert raw[off:off + 8] == struct.pack('>d', exp.reads[4][0][0, 0])


def test_fits_standard_parser_rejects_malformed_files(tmp_path):
    """The strict parser really is strict (otherwise the test above proves little)."""
    import pytest
    from tests import fits_standard as FS
    exp = _exposure(nsamp=2)
    path = exp.generate_fits(str(tmp_path))
    good = open(path, 'rb').read()
    FS.read(path)

    def broken(mutate):
        b = bytearray(good)
        mutate(b)
        p = str(tmp_path / 'bad.fits')
        open(p, 'wb').write(bytes(b))
        with pytest.raises(FS.FitsFormatError):
            FS.read(p)

    broken(lambda b: b.__setitem__(slice(len(b) - 100, len(b)), b''))            # not a multiple of 2880
    broken(lambda b: b.__setitem__(slice(80, 86), b'NAXIS '))                    # BITPIX no longer second
    broken(lambda b: b.__setitem__(slice(8, 10), b' ='))                         # value indicator
    broken(lambda b: b.__setitem__(29, ord(' ')))                                # SIMPLE's T not in column 30
    i = good.index(b'XTENSION')
    broken(lambda b: b.__setitem__(slice(i, i + 8), b'xtension'))                # lower-case keyword
