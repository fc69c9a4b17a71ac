"""Observation: the visit driver (mirror of wayne/observation.py:32-538).

Same ``setup_*`` methods, ``generate_lightcurves``, ``run_observation``,
``_generate_exposure`` and ``_generate_direct_image`` as the reference, so
``run_visit`` wires it identically.  What differs:

* the reference loop is serial (observation.py:403-405: generate, write, next);
  here exposures are queued on the device back to back and their FITS files are
  written as their device->host copies complete (``pipeline_depth`` behind), and
  a visit can be partitioned exposure-wise over GPUs (``shard=(rank, world)``,
  wayne_b200/sharding.py) -- frames do not depend on either;
* light curves come from wayne_b200/lightcurve.py (pylightcurve is not
  available): per exposure a Chebyshev planet signal evaluated inside the counts
  kernel instead of an [n_samples][n_wl] array;
* the planet is a plain :class:`Planet` record in stated units (exodata is not
  available); times are used as given (no JD -> HJD conversion: ``ephem`` is not
  available; the shipped example already quotes its transit time in HJD).
"""
from __future__ import annotations

import collections

import numpy as np

from . import lightcurve, sharding
from . import units as u
from .exposure_generator import ExposureGenerator
from .trend_generators import visit_trends
from .visit_planner import VisitPlanner

AU_IN_RSUN = 215.032
RJUP_IN_RSUN = 0.10045


class Star(object):
    def __init__(self, R=None, T=None):
        self.R = R          # solar radii
        self.T = T          # K


class Planet(object):
    """Orbital / physical parameters the driver needs (units in brackets)."""

    def __init__(self, name='None', P=None, a=None, R=None, i=None, e=0.0, periastron=0.0,
                 transittime=None, star=None):
        self.name = name
        self.P = P                      # period [days]
        self.a = a                      # semi-major axis [au]
        self.R = R                      # radius [Jupiter radii]
        self.i = i                      # inclination [deg]
        self.e = e
        self.periastron = periastron    # [deg]
        self.transittime = transittime  # mid-transit [JD, same system as the exposure times]
        self.star = star if star is not None else Star()


def detect_orbits(exp_start_times, separation=0.028):
    """Indexes at which a new orbit starts: gaps of >= ``separation`` days
    (tools.detect_orbits, wayne/tools.py:274-300)."""
    t = np.asarray(u.value_in(exp_start_times, u.day), dtype=float)
    idx = [0]
    for k in range(1, len(t)):
        if t[k] - t[k - 1] >= separation:
            idx.append(k)
    return idx


class Observation(object):
    def __init__(self, outdir=''):
        self.scanning = True
        self.outdir = outdir
        self._visit_trend = False
        self.progess = None
        self.ldcoeffs = None
        self.noise_mean = False
        self.noise_std = False
        self.rng = None                 # ExposureGenerator stream mode (None = params.rng)

    # -- set-up (same names and arguments as the reference) ----------------
    def setup_observation(self, x_ref, y_ref, spatial_scan=False, scan_speed=False):
        self.x_ref, self.y_ref = x_ref, y_ref
        self.spatial_scan, self.scan_speed = spatial_scan, scan_speed

    def setup_simulator(self, sample_rate=False, clip_values_det_limits=True, threads=2):
        self.sample_rate = sample_rate
        self.clip_values_det_limits = clip_values_det_limits
        self.threads = threads

    def setup_target(self, planet, wavelengths, planet_spectrum, stellar_flux, transittime=None,
                     ldcoeffs=None, period=None, rp=None, sma=None, inclination=None,
                     eccentricity=None, periastron=None, stellar_radius=None):
        self.planet = planet
        self.wl = wavelengths
        self.stellar_flux = stellar_flux
        self.planet_spectrum = planet_spectrum
        assert len(wavelengths) == len(stellar_flux)
        if planet_spectrum is not None:
            assert len(wavelengths) == len(planet_spectrum)
            self.transmission_spectroscopy = True
            if not isinstance(planet, Planet):
                self.planet = Planet(name=planet)
            p = self.planet
            for attr, val in (('P', period), ('R', rp), ('a', sma), ('i', inclination)):
                if val:
                    setattr(p, attr, float(u.value_in(val, None) if u.is_quantity(val) else val))
            if eccentricity or eccentricity == 0:
                p.e = eccentricity
            if periastron or periastron == 0:
                p.periastron = periastron
            if stellar_radius:
                p.star.R = float(stellar_radius)
            if transittime:
                p.transittime = float(transittime)
            if not ldcoeffs:
                raise ValueError("ldcoeffs are required (the reference looks them up with "
                                 "pylightcurve.clablimb, which is not available here)")
            self.ldcoeffs = ldcoeffs
        else:
            self.transmission_spectroscopy = False

    def setup_detector(self, detector, NSAMP, SAMPSEQ, SUBARRAY):
        self.detector, self.NSAMP, self.SAMPSEQ, self.SUBARRAY = detector, NSAMP, SAMPSEQ, SUBARRAY

    def setup_grism(self, grism):
        self.grism = grism

    def setup_visit(self, start_JD, num_orbits, exp_start_times=False):
        self.start_JD = start_JD
        self.num_orbits = num_orbits
        if exp_start_times is not False and exp_start_times is not None and len(exp_start_times):
            self.exp_start_times = exp_start_times
            self.visit_plan = {'exp_start_times': self.exp_start_times,
                               'orbit_start_index': detect_orbits(self.exp_start_times)}
        else:
            self.visit_plan = VisitPlanner(self.detector, self.NSAMP, self.SAMPSEQ, self.SUBARRAY,
                                           self.num_orbits, exp_overhead=3 * u.min)
            self.exp_start_times = self.visit_plan['exp_times'].to(u.day) + self.start_JD
            self.visit_plan['exp_start_times'] = self.exp_start_times

    def setup_reductions(self, add_dark=True, add_flat=True, add_gain_variations=True,
                         add_non_linear=True, add_initial_bias=True):
        self.add_dark, self.add_flat = add_dark, add_flat
        self.add_gain_variations, self.add_non_linear = add_gain_variations, add_non_linear
        self.add_initial_bias = add_initial_bias

    def setup_trends(self, ssv_gen, x_shifts=0, x_jitter=0.0000001, y_shifts=0, y_jitter=0.0000001):
        self.ssv_gen = ssv_gen
        self.x_shifts, self.x_jitter = x_shifts, x_jitter
        self.y_shifts, self.y_jitter = y_shifts, y_jitter

    def setup_noise_sources(self, sky_background=1 * u.count / u.s, cosmic_rate=11.,
                            add_read_noise=True, add_stellar_noise=True):
        self.sky_background = sky_background
        self.cosmic_rate = cosmic_rate
        self.add_read_noise, self.add_stellar_noise = add_read_noise, add_stellar_noise

    def setup_gaussian_noise(self, noise_mean=False, noise_std=False):
        self.noise_mean, self.noise_std = noise_mean, noise_std

    def setup_visit_trend(self, visit_trend_coeffs):
        self._visit_trend = visit_trends.HookAndLongTermRamp(self.visit_plan, visit_trend_coeffs)

    # -- light curves ---------------------------------------------------------
    def _orbit(self):
        p = self.planet
        a_rs = float(p.a) * AU_IN_RSUN / float(p.star.R)
        w = 0.0 if p.periastron is None or np.isnan(p.periastron) else float(p.periastron)
        return dict(period=float(p.P), a=a_rs, e=float(p.e or 0.0), inc_deg=float(p.i), w_deg=w,
                    t0=float(p.transittime))

    def _rp_body(self):
        """Planet radius / stellar radius used by the secondary-eclipse term (None: no term)."""
        p = self.planet
        return float(p.R) * RJUP_IN_RSUN / float(p.star.R) if (p.R and p.star.R) else None

    def generate_lightcurves(self, time_array, depth=False):
        """[len(time_array)][len(spectrum)] relative flux: transit - (1 - eclipse),
        as observation.py:293-357 (dense host evaluation; the exposure path uses
        the Chebyshev form, :meth:`_planet_signal`)."""
        t = np.asarray(u.value_in(time_array, u.day), dtype=float)
        spectrum = np.array([depth]) if depth else np.asarray(self.planet_spectrum, dtype=float)
        orb = self._orbit()
        rp_body = self._rp_body()
        models = np.zeros((len(t), len(spectrum)))
        for j, d in enumerate(spectrum):
            m = lightcurve.transit(self.ldcoeffs, np.sqrt(d), t=t, **orb)
            if rp_body:
                m = m - (1. - lightcurve.eclipse(d, rp_body, t=t, **orb))
            models[:, j] = m
        return models

    def _planet_signal(self, time_array, device=True):
        """Chebyshev planet signal of one exposure's sub-sample times, equal to
        ``1 - generate_lightcurves(time_array)`` (transit AND secondary-eclipse terms,
        observation.py:338-343, 441-443); with a CUDA device the transit quadrature
        runs on the GPU (wb200_transit_cheb)."""
        t = np.asarray(u.value_in(time_array, u.day), dtype=float)
        if device:
            import torch
            if torch.cuda.is_available():
                from .engine import DeviceEngine
                return lightcurve.planet_signal_device(DeviceEngine.get(), t, self.planet_spectrum,
                                                       self.ldcoeffs, rp_body=self._rp_body(),
                                                       **self._orbit())
        return lightcurve.planet_signal(t, self.planet_spectrum, self.ldcoeffs, rp_body=self._rp_body(),
                                        **self._orbit())

    # -- the visit ------------------------------------------------------------
    def run_observation(self, shard=None, pipeline_depth=3, write_fits=True):
        """Generate every exposure of the visit (``shard=(rank, world)``: this
        rank's exposure-wise share).  Returns {file number: path or Exposure}."""
        results = {}
        rank, world = shard if shard is not None else (0, 1)
        from . import params
        if world > 1 and (self.rng or params.rng) == 'numpy':
            # the compat mode consumes ONE sequential numpy stream across the exposures of a visit
            # (run_visit.py:73-77): ranks seeded alike would repeat each other's noise realisations
            raise ValueError("rng='numpy' (the reference's sequential streams) cannot be sharded over "
                             "{} ranks; use rng='philox'".format(world))
        if rank == 0:
            self._generate_direct_image()
        numbers = [i + 1 for i in sharding.shard_indices(len(self.exp_start_times), world, rank)]
        pending = collections.deque()

        def finish(item):
            number, filename, frame = item
            results[number] = frame.generate_fits(self.outdir, filename, ldcoeffs=self.ldcoeffs) \
                if write_fits else frame

        for number in numbers:
            start_time = self.exp_start_times[number - 1]
            pending.append(self._generate_exposure(start_time, number, write=False))
            while len(pending) > pipeline_depth:
                finish(pending.popleft())
        while pending:
            finish(pending.popleft())
        return results

    def _generate_exposure(self, expstart, number, write=True):
        index_number = number - 1
        filename = '{:04d}_raw.fits'.format(number)
        exp_gen = ExposureGenerator(self.detector, self.grism, self.NSAMP, self.SAMPSEQ, self.SUBARRAY,
                                    self.planet, filename, expstart, rng=self.rng)
        if not self.spatial_scan:
            self.sample_rate = 1 * u.year
        _, sample_mid_points, sample_durations, read_index = \
            exp_gen._gen_scanning_sample_times(self.sample_rate)
        time_array = (sample_mid_points + expstart).to(u.day)
        planet_depths = self._planet_signal(time_array) if self.transmission_spectroscopy else None

        x_ref = self._try_index(self.x_ref, index_number) + self.x_shifts * index_number
        y_ref = self._try_index(self.y_ref, index_number) + self.y_shifts * index_number
        sky_background = self._try_index(self.sky_background, index_number)
        scale_factor = self._visit_trend.get_scale_factor(index_number) if self._visit_trend else None
        common = dict(noise_mean=self.noise_mean, noise_std=self.noise_std, add_flat=self.add_flat,
                      add_dark=self.add_dark, scale_factor=scale_factor, sky_background=sky_background,
                      cosmic_rate=self.cosmic_rate, add_gain_variations=self.add_gain_variations,
                      add_non_linear=self.add_non_linear,
                      clip_values_det_limits=self.clip_values_det_limits,
                      add_read_noise=self.add_read_noise, add_stellar_noise=self.add_stellar_noise,
                      add_initial_bias=self.add_initial_bias, progress_bar=self.progess,
                      threads=self.threads)
        if self.spatial_scan:
            frame = exp_gen.scanning_frame(x_ref, y_ref, self.x_jitter, self.y_jitter, self.wl,
                                           self.stellar_flux, planet_depths, self.scan_speed,
                                           self.sample_rate, sample_mid_points, sample_durations,
                                           read_index, ssv_generator=self.ssv_gen, **common)
        else:
            frame = exp_gen.staring_frame(
                x_ref, y_ref, self.x_jitter, self.y_jitter, self.wl, self.stellar_flux, planet_depths,
                sample_mid_points, sample_durations, read_index, common['noise_mean'],
                common['noise_std'], common['add_dark'], common['add_flat'], common['cosmic_rate'],
                common['sky_background'], common['scale_factor'], common['add_gain_variations'],
                common['add_non_linear'], common['clip_values_det_limits'], common['add_read_noise'],
                common['add_stellar_noise'], common['add_initial_bias'], common['progress_bar'],
                common['threads'])
        if write:
            frame.generate_fits(self.outdir, filename, ldcoeffs=self.ldcoeffs)
            return frame
        return number, filename, frame

    @staticmethod
    def _try_index(value, index):
        try:
            return value[index]
        except (TypeError, IndexError):
            return value

    def _generate_direct_image(self):
        di_start = (self.exp_start_times[0] - 1 * u.min).to(u.day)
        gen = ExposureGenerator(self.detector, self.grism, self.NSAMP, self.SAMPSEQ, self.SUBARRAY,
                                self.planet, '0000_flt.fits', di_start)
        exp = gen.direct_image(self._try_index(self.x_ref, 0), self._try_index(self.y_ref, 0))
        exp.generate_fits(self.outdir, '0000_flt.fits')
        return exp
