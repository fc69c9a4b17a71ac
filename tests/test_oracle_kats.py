"""Pins oracle/exposure_oracle.py to every known-answer value the reference's
own tests hold for the path: trace / dispersion
(reference tests/test_grism.py:16-21, 29-51, 66-80), bin widths
(tests/test_tools.py:65-81), crop (tests/test_tools.py:12-43), the visit trend
(tests/trend_generators/test_visit_trends.py:36-51) and the SPARS10 / RAPID
read times used for sample timing (tests/test_detector.py:72-81)."""
import numpy as np
import numpy.testing as npt

from oracle import exposure_oracle as E

A, B = E.G141_TRACE, E.G141_WLSOL


def test_pixel_wl_kats():
    for args, want in (((50, 50, 100, 50), 11222.2), ((50, 50, 200, 50), 15748.6),
                       ((50, 50, 100, 51), 11222.7), ((50, 60, 100, 50), 11218.8),
                       ((60, 50, 100, 50), 10770.6)):
        assert abs(E.pixel_wl(*args, a=A, b=B) - want) < 0.05
    wl = E.pixel_wl(50, 50, np.arange(1024), 50, A, B)
    assert abs(wl.mean() - 29961.2) < 0.05 and abs(wl.min() - 8959.) < 0.05 and abs(wl.max() - 53001.1) < 0.05
    npt.assert_array_almost_equal(E.pixel_wl(50, 50, np.array([100, 110, 120, 150, 200]), 51, A, B),
                                  [11222.7, 11675.3, 12127.9, 13485.9, 15749.1], 1)
    edges = np.append(np.array([100, 110, 120, 130]) - 5., 135.)
    npt.assert_array_almost_equal(E.pixel_wl(50, 50, edges, 50, A, B),
                                  [10995.9, 11448.5, 11901.2, 12353.8, 12806.5], 1)


def test_wavelength_calibration_coeff_kats():
    npt.assert_array_almost_equal(E.wavelength_calibration_coeffs(50, 50, A, B),
                                  [0.0099, 1.8767, 45.2665, 8958.9896], 4)
    npt.assert_array_almost_equal(E.wavelength_calibration_coeffs(100, 50, A, B),
                                  [0.0096, 1.8812, 45.2776, 8963.6693], 4)
    npt.assert_array_almost_equal(E.wavelength_calibration_coeffs(50, 100, A, B),
                                  [0.0099, 1.7801, 45.3782, 8958.9896], 4)


def test_trace_worked_example():
    # SURVEY A.3 worked example, derived from grism.py:756-803 at (404.497, 457.427)
    t = E.Trace(404.497, 457.427, A, B)
    assert abs(t.m_t - 0.0089882) < 1e-7 and abs(t.c_t - 1.121572) < 1e-6
    assert abs(t.m_w - 46.27121) < 1e-5 and abs(t.c_w - 8992.168) < 1e-3
    assert abs(t.m_wl - 0.00461286) < 1e-8 and abs(t.c_wl + 0.966190) < 1e-6
    assert abs(t.wl_to_x(0.988) - 423.64) < 0.01 and abs(t.wl_to_y(0.988) - 458.72) < 0.01
    assert abs(t.wl_to_x(1.777) - 594.68) < 0.01 and abs(t.wl_to_y(1.777) - 460.26) < 0.01
    # the trace and the pixel->wavelength map agree on the trace itself
    x = t.wl_to_x(1.4)
    assert abs(E.pixel_wl(404.497, 457.427, x, t.x_to_y(x), A, B) * 1e-4 - 1.4) < 2e-3


def test_bin_widths_and_crop_kats():
    npt.assert_array_equal(E.bin_centers_to_widths([1, 2, 3, 4]), [1, 1, 1, 1])
    npt.assert_array_almost_equal(E.bin_centers_to_widths([1, 2, 4, 5.4]), [1, 1.5, 1.7, 1.4], 6)
    wl = np.arange(10.)
    for lo, hi, want in ((1, 8, np.arange(1, 9)), (0.99, 8.99, np.arange(1, 9)), (1.5, 7.5, np.arange(2, 8))):
        i0, i1 = E.crop_spectrum_ind(lo, hi, wl)
        npt.assert_array_equal(wl[i0:i1], want)


def test_visit_trend_kat():
    t = np.array([6, 9, 12, 95, 98, 101]) / 60. / 24.
    t0 = np.array([t[0]] * 3 + [t[3]] * 3)
    npt.assert_array_almost_equal(E.hook_and_long_term_ramp(t, t0, 0.005, 0.0011, 400, 9 / 60 / 24),
                                  [0.99891, 0.99952, 0.99978, 0.9986, 0.99921, 0.99947], 5)


def test_sample_times_config1():
    # SPARS10 / 256 / NSAMP 5 read times (reference tests/test_detector.py pins
    # the first two) at 10 ms: 2233 sub-samples, read_index [27, 762, 1497, 2232]
    starts, mid, dur, ri = E.gen_scanning_sample_times([0.278, 7.624, 14.971, 22.317], 10.0)
    assert len(starts) == 2233 and ri == [27, 762, 1497, 2232]
    assert abs(dur.sum() - 22317.0) < 1e-6
    assert abs(dur[27] - 8.0) < 1e-9            # a read's last sub-sample is the remainder
    npt.assert_allclose(mid, starts + dur / 2)
    # staring: one sub-sample per read interval
    starts, mid, dur, ri = E.gen_scanning_sample_times([0.278, 7.624], 365.25 * 86400e3)
    assert ri == [0, 1] and np.allclose(dur, [278., 7346.])


def test_non_linearity_roundtrip():
    cal = {'nl': [np.zeros((1024, 1024), np.float32), np.full((1024, 1024), 6.4e-7, np.float32),
                  np.zeros((1024, 1024), np.float32), np.zeros((1024, 1024), np.float32)]}
    p = np.linspace(0, 70000, 266 * 266).reshape(266, 266)
    u1, it = E.apply_non_linearity(p, cal)
    c2 = np.float32(6.4e-7)
    npt.assert_allclose(u1 * (1 + c2 * u1), p, atol=1e-4)
    assert 2 <= it < 20
