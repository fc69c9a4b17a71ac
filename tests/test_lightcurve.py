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


def test_planet_signal_carries_the_secondary_eclipse_term():
    """The reference feeds the frames planet_depths = 1 - (transit - (1 - eclipse))
    (wayne/observation.py:338-343, 441-443): during secondary eclipse the star is dimmed by
    (1 - eclipse(fp = depth[w], rp_body)) too.  The Chebyshev form the device path uses must
    equal the dense evaluation at every phase: transit, eclipse ingress / totality, and out
    of both."""
    from wayne import observation
    from wayne import units as u
    obs = observation.Observation()
    star = observation.Star(R=1.155)
    planet = observation.Planet(name='HD 209458 b', P=HD['period'], a=0.047309, R=1.38, i=HD['inc_deg'],
                                e=0.0, periastron=0.0, transittime=HD['t0'], star=star)
    wl = np.linspace(1.0, 1.7, 211)
    depth = 0.0146 * (1 + 0.02 * np.sin(9 * wl))
    obs.setup_target(planet, wl * u.micron, depth, np.ones_like(wl), ldcoeffs=LD)
    rp_body = obs._rp_body()
    assert abs(rp_body - 1.38 * 0.10045 / 1.155) < 1e-12
    mid_ecl = HD['t0'] + HD['period'] / 2
    # (at the transit's contact points the signal has a kink in rp: the order-8 expansion is good
    # to 5e-8 there, 3e-6 of the depth; the eclipse term is a smooth function of rp and exact)
    for t, tol in ((HD['t0'] + np.linspace(-0.08, 0.08, 41), 1e-7),            # transit
                   (mid_ecl + np.linspace(-0.085, -0.03, 37), 1e-12),          # eclipse ingress into totality
                   (mid_ecl + np.linspace(-0.01, 0.01, 9), 1e-12),             # totality
                   (HD['t0'] + 0.9 + np.linspace(0, 0.02, 5), 1e-15)):         # neither
        dense = 1 - obs.generate_lightcurves(t * u.day)
        cheb = obs._planet_signal(t * u.day, device=False).to_array()
        assert dense.shape == cheb.shape == (len(t), len(wl))
        assert np.abs(dense - cheb).max() < tol
        if tol == 1e-12:
            assert dense.max() > 0.01                                          # the term is really there
    tot = 1 - obs.generate_lightcurves(np.array([mid_ecl]) * u.day)[0]
    assert np.allclose(tot, depth / (1 + depth), rtol=1e-12)             # planet fully hidden
    # without a planet radius there is no eclipse term (as when rp is not configured)
    planet.R = None
    assert np.all(obs._planet_signal(np.array([mid_ecl]) * u.day, device=False).to_array() == 0)
