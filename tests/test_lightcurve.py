"""Transit model of the visit driver (wayne_b200/lightcurve.py): pinned by
properties, since pylightcurve (the reference's third-party model) is not
available here."""
import numpy as np
import pytest

from wayne_b200 import lightcurve as lc

LD = [0.800627, -0.757066, 0.897268, -0.384804]      # the shipped example's coefficients
HD = dict(period=3.524746, a=0.047309 * 215.032 / 1.155, e=0.0, inc_deg=86.71, w_deg=0.0, t0=2456196.28836)


def test_uniform_source_limit_is_geometric():
    # no limb darkening: blocked fraction = overlap area / pi
    z = np.array([0.0, 0.3, 0.85, 0.95, 1.0, 1.05, 1.2])
    p = 0.1
    f = lc.transit_flux(z, np.full(z.shape, p), [0, 0, 0, 0])
    assert np.allclose(f[:3], 1 - p * p, atol=2e-9)
    assert f[-1] == 1.0
    zz = z[3:6]
    k0 = np.arccos((p * p + zz * zz - 1) / (2 * p * zz))
    k1 = np.arccos((1 - p * p + zz * zz) / (2 * zz))
    area = p * p * k0 + k1 - 0.5 * np.sqrt(4 * zz * zz - (1 + zz * zz - p * p) ** 2)
    assert np.allclose(f[3:6], 1 - area / np.pi, atol=5e-8)


def test_limb_darkened_properties():
    t = HD['t0'] + np.linspace(-0.12, 0.12, 401)
    f = lc.transit(LD, 0.1209, t=t, **HD)
    assert f[0] == 1.0 and f[-1] == 1.0 and f.min() < 1 - 0.1209 ** 2          # deeper at the centre
    assert np.allclose(f, f[::-1], atol=1e-12)                                  # symmetric (e = 0)
    assert np.all(np.diff(f[:201]) <= 1e-12)                                    # monotonic ingress
    # quadrature converged
    f2 = lc.transit(LD, 0.1209, t=t, nodes=256, **HD)
    assert np.abs(f - f2).max() < 2e-8
    # small-planet limit: depth -> p^2 I(r) / <I>
    p = 1e-3
    c_tot = lc._claret_total(LD) / np.pi
    f0 = lc.transit_flux(np.array([0.4]), np.array([p]), LD)[0]
    assert abs((1 - f0) / (p * p) - lc._claret_intensity(np.array([0.4]), LD)[0] / c_tot) < 1e-4
    # eccentric orbit with w = 90 deg still transits at t0
    z, front = lc.kepler_separation(np.array([HD['t0']]), 3.5, 8.0, 0.3, 90.0, 90.0, HD['t0'])
    assert front[0] and z[0] < 1e-9


def test_eclipse():
    t = HD['t0'] + HD['period'] / 2 + np.linspace(-0.1, 0.1, 101)
    e = lc.eclipse(1e-3, 0.12, t=t, **HD)
    assert abs(e[0] - 1) < 1e-15 and abs(e[50] - 1 / 1.001) < 1e-12
    assert np.all(lc.eclipse(1e-3, 0.12, t=HD['t0'] + np.linspace(-0.1, 0.1, 11), **HD) == 1.0)


def test_chebyshev_signal_matches_direct_evaluation():
    depth = 0.0146 * (1 + 0.02 * np.sin(np.linspace(0, 9, 300)))
    t = HD['t0'] + np.linspace(-0.09, 0.02, 57)
    sig = lc.planet_signal(t, depth, LD, **HD)
    assert sig.shape == (57, 300) and sig.ndim == 2
    arr = sig.to_array()
    direct = np.array([1 - lc.transit(LD, np.sqrt(d), t=t, **HD) for d in depth[::37]]).T
    assert np.abs(arr[:, ::37] - direct).max() < 1e-10
    assert np.allclose(sig[5], arr[5]) and np.allclose(sig[10:20].to_array(), arr[10:20])
    assert arr.max() > 0.015 and arr.min() == 0.0


def test_linear_limb_darkening_centred_planet_closed_form():
    # I = 1 - u (1 - mu): planet centred on the disk (z = 0), blocked light in closed form
    for u_ld, p in ((0.6, 0.1), (0.3, 0.3), (1.0, 0.05)):
        blocked = np.pi * p * p * (1 - u_ld) + 2 * np.pi * u_ld / 3 * (1 - (1 - p * p) ** 1.5)
        total = np.pi * (1 - u_ld / 3)
        f = lc.transit_flux(np.array([0.0]), np.array([p]), [0.0, u_ld, 0.0, 0.0])[0]
        assert abs(f - (1 - blocked / total)) < 2e-10


@pytest.mark.filterwarnings('ignore::scipy.integrate.IntegrationWarning')
def test_against_two_dimensional_quadrature():
    """An independent integrator: the limb-darkened intensity integrated over the planet's disk
    in planet-centred Cartesian coordinates (scipy dblquad), for geometries inside the disk, on
    the limb (ingress) and grazing."""
    from scipy import integrate
    c = LD

    def inten(y, x, z):
        r2 = (x + z) ** 2 + y * y
        if r2 >= 1.0:
            return 0.0
        mu = np.sqrt(1 - r2)
        sq = np.sqrt(mu)
        return 1 - c[0] * (1 - sq) - c[1] * (1 - mu) - c[2] * (1 - mu * sq) - c[3] * (1 - mu * mu)

    total = lc._claret_total(c)
    for z, p in ((0.2, 0.12), (0.7, 0.1), (0.95, 0.1), (1.02, 0.12), (1.09, 0.1)):
        blocked, _ = integrate.dblquad(inten, -p, p, lambda x: -np.sqrt(max(p * p - x * x, 0.0)),
                                       lambda x: np.sqrt(max(p * p - x * x, 0.0)), args=(z,),
                                       epsabs=1e-10, epsrel=1e-10)
        f = lc.transit_flux(np.array([z]), np.array([p]), c)[0]
        assert abs(f - (1 - blocked / total)) < 5e-6, (z, p, f, 1 - blocked / total)
