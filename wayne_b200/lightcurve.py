"""Transit light curves for the visit driver.

The reference gets its per-sub-sample, per-wavelength transit depths from the
third-party ``pylightcurve`` (``pylc.transit('claret', ldcoeffs, rp, P, a, e, i,
W, T0, t)``, wayne/observation.py:293-357); that package is not in the reference
tree and not in this image, so the model is restated here from its published
definition -- **parity unpinned** against pylightcurve itself, pinned by
properties (tests/test_lightcurve.py: uniform-source limit, normalisation,
symmetry, convergence of the quadrature):

  * orbit: Keplerian, projected star-planet separation z(t) in stellar radii;
  * stellar disk: Claret (2000) four-coefficient law
        I(mu) = 1 - sum_n c_n (1 - mu^(n/2)),  mu = sqrt(1 - r^2);
  * flux = 1 - (intensity under the planet disk) / (total intensity), the
    occulted integral taken over annuli of the stellar disk with Gauss-Legendre
    nodes in a variable that removes the square-root behaviour at the limb.

B200-first: the exposure path needs the depth for every (sub-sample, bin) --
1.7e7 cells per exposure, a 135 MB host->device copy if it is materialised on
the host.  The depth depends on wavelength only through rp(lambda) = sqrt(depth),
smoothly, so :class:`ChebyshevSignal` carries, per sub-sample, a short Chebyshev
expansion in rp; ``k_counts`` evaluates it in place (Clenshaw) and the planet
signal never exists as an array.
"""
from __future__ import annotations

import numpy as np


def kepler_separation(t, period, a, e, inc_deg, w_deg, t0):
    """Projected separation z [stellar radii] and a flag "planet in front" for
    times t [days].  a in stellar radii, angles in degrees, t0 = mid-transit."""
    t = np.asarray(t, dtype=np.float64)
    inc = np.radians(inc_deg)
    if not e:
        phi = 2 * np.pi * (t - t0) / period
        x = a * np.sin(phi)
        y = a * np.cos(phi) * np.cos(inc)
        return np.hypot(x, y), np.cos(phi) > 0
    w = np.radians(w_deg if np.isfinite(w_deg) else 0.0)
    # true anomaly at mid-transit, then time of periastron
    f_tr = np.pi / 2 - w
    E_tr = 2 * np.arctan2(np.sqrt(1 - e) * np.sin(f_tr / 2), np.sqrt(1 + e) * np.cos(f_tr / 2))
    tp = t0 - period / (2 * np.pi) * (E_tr - e * np.sin(E_tr))
    M = 2 * np.pi * (t - tp) / period
    E = M.copy()
    for _ in range(60):
        dE = (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
        E = E - dE
        if np.max(np.abs(dE)) < 1e-14:
            break
    f = 2 * np.arctan2(np.sqrt(1 + e) * np.sin(E / 2), np.sqrt(1 - e) * np.cos(E / 2))
    r = a * (1 - e * np.cos(E))
    x = -r * np.cos(w + f)
    y = -r * np.sin(w + f) * np.cos(inc)
    z_los = r * np.sin(w + f) * np.sin(inc)
    return np.hypot(x, y), z_los > 0


def _claret_intensity(r, c):
    mu = np.sqrt(np.clip(1 - r * r, 0.0, None))
    sq = np.sqrt(mu)
    return 1 - c[0] * (1 - sq) - c[1] * (1 - mu) - c[2] * (1 - mu * sq) - c[3] * (1 - mu * mu)


def _claret_total(c):
    # int_0^1 mu^(n/2) 2 r dr = 4 / (n + 4)
    return np.pi * (1 - sum(c[n - 1] * (1 - 4.0 / (n + 4)) for n in (1, 2, 3, 4)))


_GL = {}


def _gauss_legendre(n):
    if n not in _GL:
        _GL[n] = np.polynomial.legendre.leggauss(n)
    return _GL[n]


def transit_flux(z, rp, ldcoeffs, nodes=96):
    """Relative flux for separations z [...] and radius ratios rp [...] (broadcast
    together).  1 outside transit."""
    z, rp = np.broadcast_arrays(np.asarray(z, dtype=np.float64), np.asarray(rp, dtype=np.float64))
    c = [float(v) for v in ldcoeffs]
    out = np.ones(z.shape)
    hit = (z < 1 + rp) & (rp > 0)
    if not hit.any():
        return out
    zz, pp = z[hit], rp[hit]
    lo = np.clip(zz - pp, 0.0, 1.0)
    hi = np.clip(zz + pp, 0.0, 1.0)
    # r = hi - (hi - lo) s^2 puts the limb's square-root behaviour at s -> 0 where
    # the substitution's Jacobian 2 s vanishes: smooth integrand for Gauss-Legendre
    x, wgt = _gauss_legendre(nodes)
    s = 0.5 * (x + 1)[:, None]
    w = 0.5 * wgt[:, None]
    span = (hi - lo)[None, :]
    r = hi[None, :] - span * s * s
    jac = 2 * span * s
    with np.errstate(invalid='ignore', divide='ignore'):
        cosang = (r * r + zz[None, :] ** 2 - pp[None, :] ** 2) / (2 * r * zz[None, :])
    ang = np.arccos(np.clip(cosang, -1.0, 1.0))
    # annuli entirely inside the planet disk (r <= p - z) are fully covered
    ang = np.where(r <= (pp - zz)[None, :], np.pi, ang)
    ang = np.where(zz[None, :] == 0, np.where(r <= pp[None, :], np.pi, 0.0), ang)
    blocked = (w * jac * _claret_intensity(r, c) * 2 * r * ang).sum(axis=0)
    out[hit] = 1 - blocked / _claret_total(c)
    return out


def transit(ldcoeffs, rp, period, a, e, inc_deg, w_deg, t0, t, nodes=96):
    """Claret-law transit light curve at times t for one radius ratio rp
    (signature order of pylightcurve's transit() minus the law name)."""
    z, front = kepler_separation(t, period, a, e, inc_deg, w_deg, t0)
    z = np.where(front, z, np.inf)
    return transit_flux(z, np.full(z.shape, float(rp)), ldcoeffs, nodes)


def eclipse_visibility(rp, period, a, e, inc_deg, w_deg, t0, t):
    """Fraction of the planet's disk (radius ratio rp) NOT hidden behind the star at times t."""
    z, front = kepler_separation(t, period, a, e, inc_deg, w_deg, t0)
    z = np.where(front, np.inf, z)
    p = float(rp)
    vis = np.ones(z.shape)
    full = z <= 1 - p
    part = (z < 1 + p) & ~full
    vis[full] = 0.0
    if part.any():
        zz = z[part]
        k0 = np.arccos(np.clip((p * p + zz * zz - 1) / (2 * p * zz), -1, 1))
        k1 = np.arccos(np.clip((1 - p * p + zz * zz) / (2 * zz), -1, 1))
        area = p * p * k0 + k1 - 0.5 * np.sqrt(np.clip(4 * zz * zz - (1 + zz * zz - p * p) ** 2, 0, None))
        vis[part] = 1 - area / (np.pi * p * p)
    return vis


def eclipse(fp_over_fs, rp, period, a, e, inc_deg, w_deg, t0, t):
    """Secondary eclipse of a uniformly bright planet: (1 + fp * visible) / (1 + fp)."""
    vis = eclipse_visibility(rp, period, a, e, inc_deg, w_deg, t0, t)
    return (1 + fp_over_fs * vis) / (1 + fp_over_fs)


def _cheb_nodes(order, pmin, pmax):
    k = np.arange(order)
    xk = np.cos(np.pi * (k + 0.5) / order)
    return k, 0.5 * (pmax - pmin) * xk + 0.5 * (pmax + pmin)


def _cheb_transform(f, order):
    """Chebyshev coefficients [..., order] of node values f [..., order]."""
    k = np.arange(order)
    T = np.cos(np.pi * np.outer(k, k + 0.5) / order)                       # T_j(x_k)
    coef = (2.0 / order) * f @ T.T
    coef[..., 0] *= 0.5
    return coef


def eclipse_cheb_term(t, rp_body, order, pmin, pmax, period, a, e, inc_deg, w_deg, t0):
    """The secondary-eclipse part of the reference's planet signal as Chebyshev
    coefficients [n_samples][order], or None when no sample is in eclipse.

    The reference dims the star by 1 - eclipse(fp = depth[w], rp_body, t) as well
    (wayne/observation.py:338-343, 441-443: planet_depths = 1 - (transit - (1 -
    eclipse))).  With a uniformly bright planet that is
        (1 - visible(t)) * fp / (1 + fp),   fp = rp[w]^2,
    a product of a per-sample factor and a smooth function of the bin's radius
    ratio -- the same expansion variable as the transit term."""
    hidden = 1.0 - eclipse_visibility(rp_body, period, a, e, inc_deg, w_deg, t0, t)
    if not np.any(hidden > 0):
        return None
    _, pk = _cheb_nodes(order, pmin, pmax)
    g = _cheb_transform((pk * pk / (1 + pk * pk))[None, :], order)[0]
    return hidden[:, None] * g[None, :]


def planet_signal_device(engine, t, depth_spectrum, ldcoeffs, period, a, e, inc_deg, w_deg, t0, order=8,
                         nodes=96, rp_body=None):
    """:func:`planet_signal` with the quadrature on the GPU (wb200_transit_cheb):
    the coefficients stay in HBM and go straight into k_counts.  The host only
    solves Kepler's equation for the sub-sample times (and adds the secondary-
    eclipse term on the rare exposures that see one)."""
    import ctypes as C

    from . import _lib
    rp = np.sqrt(np.asarray(depth_spectrum, dtype=np.float64))
    pmin, pmax = float(rp.min()), float(rp.max())
    if pmax - pmin < 1e-12:
        pmax = pmin + 1e-12
    z, front = kepler_separation(t, period, a, e, inc_deg, w_deg, t0)
    z = np.where(front, z, 1e30)
    glx, glw = engine.cached_plane(('gauss_legendre', nodes),
                                   lambda: engine.to_dev_many(list(_gauss_legendre(nodes))))
    d_z, = engine.to_dev_many([z], ahead=True)
    coef = engine.empty((len(z), order))
    ld = np.ascontiguousarray(ldcoeffs, dtype=np.float64)
    _lib.check(_lib.lib.wb200_transit_cheb(len(z), order, C.c_void_p(d_z.data_ptr()), pmin, pmax,
                                           ld.ctypes.data_as(_lib.DP), nodes, C.c_void_p(glx.data_ptr()),
                                           C.c_void_p(glw.data_ptr()), C.c_void_p(coef.data_ptr()),
                                           engine.stream_ptr()), "wb200_transit_cheb")
    if rp_body:
        ecl = eclipse_cheb_term(t, rp_body, order, pmin, pmax, period, a, e, inc_deg, w_deg, t0)
        if ecl is not None:
            d_ecl, = engine.to_dev_many([ecl], ahead=True)
            coef += d_ecl
    x = (2 * rp - (pmax + pmin)) / (pmax - pmin)
    return ChebyshevSignal(coef, x)


class ChebyshevSignal(object):
    """Planet signal (1 - relative flux) of an exposure as a per-sub-sample
    Chebyshev expansion in the radius ratio:  depth[s][w] = sum_k coef[s][k] T_k(x[w]),
    x[w] = (2 rp[w] - (pmin + pmax)) / (pmax - pmin).  ``to_array()`` materialises
    [n_samples][n_wl] on the host (tests, compat mode); the device path hands
    ``coef`` and ``x`` to k_counts."""

    ndim = 2

    def __init__(self, coef, x):
        # coef may be a CUDA tensor (planet_signal_device): it then never leaves HBM
        self.coef = coef if hasattr(coef, 'is_cuda') else np.ascontiguousarray(coef, dtype=np.float64)
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.shape = (self.coef.shape[0], self.x.shape[0])

    def _host_coef(self):
        return self.coef.cpu().numpy() if hasattr(self.coef, 'is_cuda') else self.coef

    def __getitem__(self, item):
        if isinstance(item, slice):
            return ChebyshevSignal(self.coef[item], self.x)
        return np.polynomial.chebyshev.chebval(self.x, self._host_coef()[item])

    def to_array(self):
        return np.polynomial.chebyshev.chebval(self.x, self._host_coef().T)


class SeparableSignal(object):
    """Planet signal of an exposure that factorises: depth[s][w] = lightcurve[s] * depth[w]
    (one light-curve shape times a depth spectrum -- what a caller builds with
    ``depth[None, :] * lightcurve[:, None]`` before handing it to scanning_frame).  Passed
    as this object, the two factors are uploaded (8 (N + W) bytes instead of 8 N W: 65 KB
    instead of 135 MB for a 4116 x 4096 exposure) and the counts kernel forms the very same
    double product per cell, so the exposure is bit-identical to the dense array's."""

    ndim = 2

    def __init__(self, lightcurve, depth):
        self.row = np.ascontiguousarray(lightcurve, dtype=np.float64)
        self.col = np.ascontiguousarray(depth, dtype=np.float64)
        if self.row.ndim != 1 or self.col.ndim != 1:
            raise ValueError("lightcurve [n_samples] and depth [n_wl] must be one-dimensional")
        self.shape = (self.row.shape[0], self.col.shape[0])

    def __getitem__(self, item):
        if isinstance(item, slice):
            return SeparableSignal(self.row[item], self.col)
        return self.col * self.row[item]

    def to_array(self):
        return self.col[None, :] * self.row[:, None]


def planet_signal(t, depth_spectrum, ldcoeffs, period, a, e, inc_deg, w_deg, t0, order=8, nodes=96,
                  rp_body=None):
    """ChebyshevSignal of the reference's planet signal for sample times t [days]
    and a transit-depth spectrum: ``1 - star_norm_flux`` of wayne/observation.py:
    441-443 with star_norm_flux = transit(rp = sqrt(depth[w])) - (1 - eclipse(fp =
    depth[w], rp_body)) (:338-343).  ``rp_body`` = the planet's radius ratio used
    for the secondary eclipse (None: no eclipse term, as when the planet has no
    radius)."""
    rp = np.sqrt(np.asarray(depth_spectrum, dtype=np.float64))
    pmin, pmax = float(rp.min()), float(rp.max())
    if pmax - pmin < 1e-12:
        pmax = pmin + 1e-12
    _, pk = _cheb_nodes(order, pmin, pmax)
    z, front = kepler_separation(t, period, a, e, inc_deg, w_deg, t0)
    z = np.where(front, z, np.inf)
    f = 1 - transit_flux(z[:, None], pk[None, :], ldcoeffs, nodes)        # [N][order]
    coef = _cheb_transform(f, order)
    if rp_body:
        ecl = eclipse_cheb_term(t, rp_body, order, pmin, pmax, period, a, e, inc_deg, w_deg, t0)
        if ecl is not None:
            coef = coef + ecl
    x = (2 * rp - (pmax + pmin)) / (pmax - pmin)
    return ChebyshevSignal(coef, x)
