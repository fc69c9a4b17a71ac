"""Parity AT THE SHAPE THE METRIC IS QUOTED ON (BASELINE configs[3]; bench.py's
default workload `c4`): G141 spatial scan, SUBARRAY 1024 (F = 1024, L = 1014, the
-5 flat / sub-array offset), NSAMP 15 RAPID (R = 14 read intervals), 10 ms
sub-samples (N = 4116), W = 4096 bins -- the geometry of wayne/exposure_generator.py:
336-394 that smaller cases do not reach: bin chunking at 1376 / 2048 bins per CTA, the
read-interval assignment over 14 planes, the int64 fixed-point planes near 1e9
electrons.

 (a) compat mode at low flux against oracle.exposure_oracle.scanning_frame:
     all 15 reads, every term on;
 (b) native mode at 1e9 electrons with the read-stage terms off: exact electron
     conservation per read interval, then every pixel against the analytic
     expectation (tests/analytic.py, pinned to the reference thrower on the CPU)
     within Poisson scatter;
 (c) the same bookkeeping the bench asserts in its timed loop.
"""
import numpy as np
import pytest

import bench
from oracle import exposure_oracle as E
from tests import analytic, harness

pytestmark = pytest.mark.gpu


def _workload(calb_dir, photons):
    from wayne_b200 import calibration, params
    wk = dict(bench.WORKLOADS['c4'], photons=photons)
    calibration.write_synthetic_calibration(calb_dir, modes=((wk['sub'], wk['seq']),))
    params.set_calibration_dir(calb_dir)
    return wk, bench.make_inputs(wk)


def test_c4_compat_matches_oracle(calb_dir):
    """(a) 4116 sub-samples x 4096 bins, ~2e7 electrons, rand_r streams, every
    reduction on: the 15 reads equal the oracle's (<= 1e-9 of the peak; the zero
    read bit for bit) and the electron count is identical."""
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    wk, inp = _workload(calb_dir, 2.0e7)
    cal = harness.oracle_calibration('G141', dark_mode=(1024, 'RAPID'), nsamp=15)
    kw = bench.frame_kwargs(wk, 0)
    mid = np.asarray(u.value_in(inp['mid'], u.ms))
    dur = np.asarray(u.value_in(inp['dur'], u.ms))
    assert len(mid) == 4116 and len(inp['wl']) == 4096 and len(inp['read_index']) == 14
    depth = inp['depth0'][None, :] * inp['lightcurve'][:, None]
    eg = ExposureGenerator(*inp['eg_args'], rng='numpy')
    np.random.seed(1963)
    exp = eg.scanning_frame(kw['x_ref'], kw['y_ref'], 0.025, 0.025, inp['wl'] * u.micron, inp['flux'], depth,
                            kw['scan_speed'], kw['sample_rate'], inp['mid'], inp['dur'], inp['read_index'],
                            ssv_generator=kw['ssv_generator'], cosmic_rate=11., sky_background=kw['sky_background'],
                            scale_factor=kw['scale_factor'], threads=2)
    o = E.scanning_frame(cal, 'G141', 1024, inp['read_times'], inp['wl'], inp['flux'], depth, kw['x_ref'],
                         kw['y_ref'], 0.025, 0.025, wk['scan'] * 0.001, wk['rate'], np.random.RandomState(1963),
                         ssv=(1.5, 1.1, 0), cosmic_rate=11., sky_background=5.5, scale_factor=kw['scale_factor'],
                         threads=2, sample_times=(mid, dur, inp['read_index']))
    assert eg.photons == o['photons'] and 1.5e7 < o['photons'] < 2.5e7
    assert eg._run.win_geometry[2] == 1376               # the bench's bins-per-CTA chunking
    peak = max(np.abs(r).max() for r in o['reads'])
    assert len(exp.reads) == 15
    for r, want in enumerate(o['reads']):
        got = exp.reads[r][0]
        assert got.shape == (1024, 1024)
        assert np.max(np.abs(got - want)) <= 1e-9 * peak, r
    assert np.array_equal(exp.reads[0][0], o['reads'][0])
    assert o['newton_iters'] == [int(x) for x in eg._run.newton_iters.cpu().numpy()[:14]]


def _native(inp, wk, key, add_flat):
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    kw = bench.frame_kwargs(wk, 0)
    eg = ExposureGenerator(*inp['eg_args'], rng='philox')
    depth = inp['depth0'][None, :] * inp['lightcurve'][:, None]
    exp = eg.scanning_frame(kw['x_ref'], kw['y_ref'], 0.025, 0.025, inp['wl'] * u.micron, inp['flux'], depth,
                            kw['scan_speed'], kw['sample_rate'], inp['mid'], inp['dur'], inp['read_index'],
                            ssv_generator=kw['ssv_generator'], scale_factor=kw['scale_factor'],
                            add_dark=False, add_flat=add_flat, cosmic_rate=None, sky_background=0 * u.count / u.s,
                            add_gain_variations=False, add_non_linear=False, clip_values_det_limits=False,
                            add_read_noise=False, rng_key=key)
    reads = np.array([r[0] for r in exp.reads])
    planes = np.diff(reads, axis=0) * 2.35            # electrons (x flat) per read interval, bordered
    return eg, planes, kw


def test_c4_native_conserves_every_electron(calb_dir):
    """(b1) 1e9 electrons, flat off: each read interval's plane holds exactly the
    electrons drawn for that interval's sub-samples (this pointing keeps the whole
    PSF inside the frame), as integers per pixel; the kernel's own tally agrees."""
    wk, inp = _workload(calb_dir, 1.0e9)
    eg, planes, _ = _native(inp, wk, (1963, 77), add_flat=False)
    run = eg._run
    counts = run.counts_host().astype(np.int64)
    assert counts.shape == (4116, 4096) and 9.7e8 < counts.sum() < 1.03e9
    assert run.win_geometry[2] == 2048
    ints = np.rint(planes)
    assert np.max(np.abs(planes - ints)) < 1e-6         # whole electrons in every pixel
    first = 0
    for r, last in enumerate(inp['read_index']):
        assert int(ints[r].sum()) == int(counts[first:last + 1].sum()), r
        first = last + 1
    # nothing in the reference border, nothing in row / column 0 of the light-sensitive area
    assert ints[:, :5].sum() == 0 and ints[:, :, :5].sum() == 0 and ints[:, 5].sum() == 0 and ints[:, :, 5].sum() == 0
    tally = run.tally.cpu().numpy()
    assert int(tally[0]) == int(counts.sum()) == eg.photons and int(tally[1]) == 0


def test_c4_native_pixels_match_analytic_expectation(calb_dir):
    """(b2) 1e9 electrons, flat on: every pixel of every read interval against
    sum_{sub-sample, bin} counts x exact double-Gaussian cell probability x that
    sub-sample's flat value (float64, independent code), within Poisson scatter."""
    from wayne import units as u
    wk, inp = _workload(calb_dir, 1.0e9)
    eg, planes, kw = _native(inp, wk, (1963, 78), add_flat=True)
    run = eg._run
    cal = harness.oracle_calibration('G141', dark_mode=None)
    counts = run.counts_host()
    # the exposure's own sub-sample reference positions, restated: native-mode jitter stream
    # (exposure_generator.py) + scan, then the oracle's trace for every sub-sample
    key = (1963, 78)
    g = np.random.Generator(np.random.Philox(key=(key[0] << 32) | key[1]))
    N = counts.shape[0]
    jx, jy = g.normal(0, 1, N) * 0.025, g.normal(0, 1, N) * 0.025
    mid = np.asarray(u.value_in(inp['mid'], u.ms))
    xr = kw['x_ref'] + jx
    yr = kw['y_ref'] + mid * (wk['scan'] * 0.001) + jy
    i0, i1 = E.crop_spectrum_ind(*E.WL_LIMITS['G141'], inp['wl'])
    s_wl = inp['wl'][i0:i1]
    ratio, sigl, sigh, _, _ = E.bin_tables(s_wl, cal['sens_wl_um'], cal['sens_val'])
    tr = E.Trace(xr[:, None], yr[:, None], E.G141_TRACE, E.G141_WLSOL)
    sub_scale = 507 - 1024 // 2
    xs = tr.wl_to_x(s_wl[None, :]) - sub_scale
    ys = tr.wl_to_y(s_wl[None, :]) - sub_scale
    assert np.max(np.abs(run.trace_host()[:, 0] - xr)) < 1e-9
    exp, var = analytic.expected_interval_images(counts, xs, ys, ratio, sigl, sigh, inp['read_index'], 1014,
                                                 cal=cal, grism_name='G141', subarray=1024,
                                                 refs=np.stack([xr, yr], axis=1), with_var=True)
    got = planes[:, 5:-5, 5:-5]
    assert got.shape == exp.shape == (14, 1014, 1014)
    # totals: flat-weighted electrons per interval
    for r in range(14):
        assert abs(got[r].sum() - exp[r].sum()) < 6 * np.sqrt(exp[r].sum()), r
    m = exp > 50
    assert m.sum() > 1.5e5
    # the electrons of a bin are multinomial over the pixels: scatter measured against the exact variance
    z = (got[m] - exp[m]) / np.sqrt(var[m])
    assert abs(z.mean()) < 6 / np.sqrt(m.sum())
    assert 0.985 < z.std() < 1.015
    assert np.abs(z).max() < 6.5
    # faint halo: where less than one electron is expected per pixel, the totals still agree
    halo = (exp < 1.0) & (exp > 0)
    assert abs(got[halo].sum() - exp[halo].sum()) < 6 * np.sqrt(exp[halo].sum() + 1)
    assert got[exp == 0].sum() == 0
    # per read interval and per detector row (the scan direction): no interval is shifted
    for r in (0, 6, 13):
        rows_g, rows_e = got[r].sum(axis=1), exp[r].sum(axis=1)
        mm = rows_e > 1e4
        zr = (rows_g[mm] - rows_e[mm]) / np.sqrt(rows_e[mm])
        assert np.abs(zr).max() < 5.5, r


def test_c4_bench_bookkeeping(calb_dir):
    """(c) what bench.py asserts after its timed loop: binned + dropped == thrown."""
    wk, inp = _workload(calb_dir, 1.0e8)
    eg, planes, _ = _native(inp, wk, (1963, 79), add_flat=True)
    tally = eg._run.tally.cpu().numpy()
    assert int(tally[0]) + int(tally[1]) == eg.photons > 9e7
