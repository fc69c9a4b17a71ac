"""Host-side mirror of the reference interface, tested the way the reference
tests its own (tests/test_grism.py, test_detector.py, test_tools.py,
trend_generators/test_*.py): same calls through the ``wayne`` alias, same
expected numbers."""
import numpy as np
import numpy.testing as npt
import pytest

from wayne import detector, grism, tools
from wayne import units as u
from wayne.trend_generators import cosmic_rays, scan_speed_varations, visit_trends


class Test_G141_Grism:
    def setup_method(self):
        self.g = grism.G141()

    def test_get_pixel_wl(self):
        for args, want in (((50, 50, 100, 50), 11222.2), ((50, 50, 200, 50), 15748.6),
                           ((50, 50, 100, 51), 11222.7), ((50, 60, 100, 50), 11218.8),
                           ((60, 50, 100, 50), 10770.6)):
            assert abs(self.g.get_pixel_wl(*args) - want) < 0.05

    def test_get_pixel_wl_per_row(self):
        wl = self.g.get_pixel_wl_per_row(50, 50, np.arange(1024))
        assert len(wl) == 1024
        assert abs(wl.mean() - 29961.2) < 0.05 and abs(wl.min() - 8959.) < 0.05
        assert abs(wl.max() - 53001.1) < 0.05
        npt.assert_array_almost_equal(
            self.g.get_pixel_wl_per_row(50, 50, np.array([100, 110, 120, 150, 200])),
            [11222.2, 11674.8, 12127.5, 13485.4, 15748.6], 1)
        npt.assert_array_almost_equal(
            self.g.get_pixel_wl_per_row(50, 50, np.array([100, 110, 120, 150, 200]), 51),
            [11222.7, 11675.3, 12127.9, 13485.9, 15749.1], 1)

    def test_get_pixel_edges_wl_per_row(self):
        wl = self.g.get_pixel_edges_wl_per_row(50, 50, np.array([100, 110, 120, 130]), None, 10)
        npt.assert_array_almost_equal(wl, [10995.9, 11448.5, 11901.2, 12353.8, 12806.5], 1)

    def test_bin_centers_to_limits(self):
        npt.assert_array_equal(self.g._bin_centers_to_limits(np.array([-1, 0, 1]), 1), np.arange(-1.5, 2.))

    def test_trace_coeffs(self):
        grism._SpectrumTrace(50, 50, np.zeros(9) + 1e-9, np.zeros(9) + 1e-9)
        t = grism.G141_Trace(50, 50)
        npt.assert_array_almost_equal(np.array(t._get_wavelength_calibration_coeffs(50, 50)),
                                      [0.0099, 1.8767, 45.2665, 8958.9896], 4)
        npt.assert_array_almost_equal(np.array(t._get_wavelength_calibration_coeffs(100, 50)),
                                      [0.0096, 1.8812, 45.2776, 8963.6693], 4)
        npt.assert_array_almost_equal(np.array(t._get_wavelength_calibration_coeffs(50, 100)),
                                      [0.0099, 1.7801, 45.3782, 8958.9896], 4)

    def test_g102(self):
        g = grism.G102()
        assert g.name == 'G102' and g.flat_file_name == grism.G141.FLAT_FILE   # SURVEY B3, faithful
        assert grism.G102(use_own_flat=True).flat_file_name.startswith('WFC3.IR.G102')
        assert float(u.value_in(g.wl_limits[0], u.micron)) == 0.75


class Test_WFC3_IR:
    def test_get_modes(self):
        df, df2 = detector.WFC3_IR()._get_modes()
        assert df.shape == (360, 4) and df2.shape == (19, 3)

    def test_get_exptime_works(self):
        det = detector.WFC3_IR()
        assert det.exptime(NSAMP=2, SAMPSEQ='RAPID', SUBARRAY=1024) == 2.932 * u.s
        assert det.exptime(NSAMP=16, SAMPSEQ='RAPID', SUBARRAY=64) == 0.912 * u.s
        # the reference's assertion has no abs() (tests/test_detector.py:32); the table row
        # for NSAMP=8 (SAMPNUM 7) is 138.381 s
        assert det.exptime(NSAMP=8, SAMPSEQ='SPARS25', SUBARRAY=512).value - 161.302 < 0.001
        assert det.exptime(NSAMP=8, SAMPSEQ='SPARS25', SUBARRAY=512).value == 138.381
        x = 4. * u.s
        x += det.exptime(NSAMP=2, SAMPSEQ='RAPID', SUBARRAY=1024) + 6. * u.s

    @pytest.mark.parametrize("kw", [dict(NSAMP=17, SAMPSEQ='RAPID', SUBARRAY=1024),
                                    dict(NSAMP=0, SAMPSEQ='RAPID', SUBARRAY=1024),
                                    dict(NSAMP=15, SAMPSEQ='WRONG', SUBARRAY=1024),
                                    dict(NSAMP=15, SAMPSEQ='SPARS25', SUBARRAY=128),
                                    dict(NSAMP=15, SAMPSEQ='RAPID', SUBARRAY=1023),
                                    dict(NSAMP=15, SAMPSEQ='RAPID', SUBARRAY=0)])
    def test_invalid_modes_raise(self, kw):
        det = detector.WFC3_IR()
        with pytest.raises(detector.WFC3SimSampleModeError):
            det.exptime(**kw)
        with pytest.raises(detector.WFC3SimSampleModeError):
            det.get_read_times(**kw)

    def test_get_read_times_works(self):
        det = detector.WFC3_IR()
        npt.assert_array_almost_equal(
            det.get_read_times(NSAMP=5, SAMPSEQ='RAPID', SUBARRAY=1024).to(u.s).value,
            [2.932, 5.865, 8.797, 11.729], 3)
        npt.assert_array_almost_equal(
            det.get_read_times(NSAMP=3, SAMPSEQ='SPARS10', SUBARRAY=256).to(u.s).value,
            [0.278, 7.624], 3)

    def test_geometry(self):
        det = detector.WFC3_IR()
        assert det.gen_pixel_array(1024).shape == (1014, 1014)
        assert det.gen_pixel_array(1024, light_sensitive=False).shape == (1024, 1024)
        assert det.gen_pixel_array(256, light_sensitive=False).shape == (266, 266)
        full = det.add_bias_pixels(np.ones((256, 256)))
        assert full.shape == (266, 266) and full.sum() == 256 * 256 and full[4, 5] == 0
        with pytest.raises(ValueError):
            det.add_bias_pixels(np.ones((100, 100)))
        assert det.get_initial_bias().shape == (266, 266)
        assert det.num_exp_per_buffer(5, 256) == 21


def test_tools_kats():
    wl = np.arange(10.)
    flux = wl * 2
    for lo, hi, want in ((1, 8, np.arange(1, 9)), (0.99, 8.99, np.arange(1, 9)), (1.5, 7.5, np.arange(2, 8))):
        cw, cf = tools.crop_spectrum(lo, hi, wl, flux)
        npt.assert_array_equal(cw, want)
        npt.assert_array_equal(cf, want * 2)
    npt.assert_array_equal(tools.bin_centers_to_edges(np.array([1, 2, 3, 4])), [0.5, 1.5, 2.5, 3.5, 4.5])
    npt.assert_array_almost_equal(tools.bin_centers_to_edges(np.array([1, 2, 4, 5.4])), [0.5, 1.5, 3, 4.7, 6.1], 6)
    npt.assert_array_equal(tools.bin_centers_to_widths(np.array([1, 2, 3, 4])), [1, 1, 1, 1])
    npt.assert_array_almost_equal(tools.bin_centers_to_widths(np.array([1, 2, 4, 5.4])), [1, 1.5, 1.7, 1.4], 6)
    a = np.arange(36.).reshape(6, 6)
    npt.assert_array_equal(tools.crop_central_box(a, 2), a[2:4, 2:4])
    assert tools.crop_central_box(a, 6) is a and tools.crop_central_box(a, 8) is a


def test_visit_trend_kat():
    plan = {'exp_start_times': (np.array([6, 9, 12, 95, 98, 101]) * u.min).to(u.day),
            'orbit_start_index': [0, 3]}
    vt = visit_trends.HookAndLongTermRamp(plan, (0.005, 0.0011, 400, 9 / 60 / 24))
    npt.assert_array_almost_equal(vt.scale_factors,
                                  [0.99891, 0.99952, 0.99978, 0.9986, 0.99921, 0.99947], 5)
    assert vt.get_scale_factor(3) == vt.scale_factors[3]


def test_cosmic_generators():
    g = cosmic_rays.BaseCosmicGenerator()
    assert g._number_of_cosmics(1) == 11 and g._number_of_cosmics(2) == 22
    assert g._generate_cosmic_energies(2) == [25000, 25000]
    assert g._generate_array((20, 30)).shape == (20, 30)
    arr = g._cosmics_to_array([1, 10, 5], np.zeros((10, 10)))
    assert arr.sum() == 16
    assert g.cosmic_frame(2, 10).sum() == 11 * 2 * 25000
    m = cosmic_rays.MinMaxPossionCosmicGenerator(11, 10000, 35000)
    np.random.seed(4)
    assert 10 <= np.mean([m._number_of_cosmics(1) for _ in range(200)]) <= 12
    en = m._generate_cosmic_energies(3)
    assert len(en) == 3 and all(10000 <= e <= 35000 for e in en)
    # sparse hit list == dense frame for the same numpy state
    np.random.seed(9)
    frame = m.cosmic_frame(20., 64)
    np.random.seed(9)
    rows, cols, energies = m.cosmic_hits(20., 64)
    dense = np.zeros((64, 64))
    np.add.at(dense, (rows, cols), energies)
    npt.assert_array_equal(frame, dense)


def test_ssv_sine():
    y = np.linspace(100., 260., 50)
    d = np.full(50, 10.) * u.ms
    out = scan_speed_varations.SSVSine(1.5, 1.1, 0).get_subsample_exposure_times(y, d, None, None)
    npt.assert_allclose(u.value_in(out, u.ms), 10. * (1 + 0.015 * np.sin(1.1 * (y - y[0]))))


def test_units_subset():
    assert abs((7.4325 * u.pixel / u.s).to(u.pixel / u.ms).value - 0.0074325) < 1e-15
    assert abs((1 * u.year).to(u.ms).value - 365.25 * 86400e3) < 1
    assert ((10 * u.ms) * (2 * u.pixel / u.ms)).to(u.pixel).value == 20
    with pytest.raises(Exception):
        (1 * u.s).to(u.pixel)
