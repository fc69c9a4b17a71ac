"""ExposureGenerator: per-exposure detector-image synthesis on one B200.

Host-side mirror of ``wayne.exposure_generator.ExposureGenerator``
(wayne/exposure_generator.py:16-727): same constructor, ``direct_image``,
``staring_frame`` (all-positional), ``scanning_frame`` (same keywords and
defaults), ``_gen_scanning_sample_times`` (called by Observation,
wayne/observation.py:436-437), attributes ``exptime``, ``read_times``,
``exp_info``, ``exposure``; returns ``exposure.Exposure`` with ``reads`` =
``[(ndarray, header)]``, zero read first.

What differs is the mechanism.  The reference loops over sub-samples in Python
(:336-394), calling the C electron thrower once per sub-sample and making
~12 numpy passes per read.  Here one exposure is one batched device job
(engine.ExposureRun): stage-1 tables and traces, counts, one photon launch per
window batch, the ordered flat-field gather and one fused per-pixel pass over
the whole ramp, all on one CUDA stream with a single device->host copy at the end.
There is no CPU implementation of the path in this package.

Random streams (``rng``, default ``params.rng``):
  'philox'  native.  Every draw comes from Philox4x32-10 keyed by ``rng_key``
            (default: visit seed, CRC of the file name) -- independent of launch
            geometry, batching and of how exposures are spread over GPUs.
  'numpy'   compat.  The numpy global RandomState is consumed in exactly the
            reference's order (seeds, jitter, per-sub-sample Poisson,
            per-read noise / sky / cosmics, dark, read noise) and the electrons
            come from the reference's rand_r streams reproduced on the GPU for
            the given ``threads`` -- the mode the bit-exact parity tests use.
"""
from __future__ import annotations

import time
import warnings
import zlib

import numpy as np

from . import exposure, filters, params, tools
from . import units as u
from .detector import WFC3SimNoDarkFileError
from .trend_generators import cosmic_rays, scan_speed_varations

BORDER = 5


class WFC3SimNoDarkFileWarning(Warning):
    pass


def _ms(x):
    return np.asarray(u.value_in(x, u.ms), dtype=np.float64)


class ExposureGenerator(object):
    def __init__(self, detector, grism, NSAMP, SAMPSEQ, SUBARRAY, planet,
                 filename='0001_raw.fits', start_JD=0 * u.day, rng=None, device=None):
        self.detector = detector
        self.grism = grism
        self.planet = planet
        self.NSAMP = NSAMP
        self.SAMPSEQ = SAMPSEQ
        self.SUBARRAY = SUBARRAY
        self.rng = rng
        self.device = device

        self.exptime = self.detector.exptime(NSAMP, SUBARRAY, SAMPSEQ)
        self.read_times = self.detector.get_read_times(NSAMP, SUBARRAY, SAMPSEQ)

        self.exp_info = {
            'filename': filename,
            'EXPSTART': start_JD,
            'EXPEND': start_JD + self.exptime.to(u.day),
            'EXPTIME': self.exptime.to(u.s),
            'SCAN': False,
            'SCAN_DIR': None,
            'OBSTYPE': 'SPECTROSCOPIC',
            'NSAMP': self.NSAMP,
            'SAMPSEQ': self.SAMPSEQ,
            'SUBARRAY': self.SUBARRAY,
            'samp_rate': 0 * u.s,
            'sim_time': 0 * u.s,
            'scan_speed_var': False,
            'noise_mean': False,
            'noise_std': False,
            'add_dark': False,
            'add_stellar_noise': False,
        }

    # ------------------------------------------------------------------
    def direct_image(self, x_ref, y_ref):
        """Unscaled 2-D Gaussian at the source position (zero read + one read);
        host-side numpy, as in the reference (exposure_generator.py:83-144)."""
        self.exp_info.update({
            'OBSTYPE': 'IMAGING', 'x_ref': x_ref, 'NSAMP': 2, 'SAMP-SEQ': 'RAPID', 'y_ref': y_ref,
            'add_flat': False, 'add_gain': False, 'add_non_linear': False, 'add_read_noise': False,
            'cosmic_rate': 0, 'sky_background': 0 * u.ct / u.s, 'scale_factor': 1,
            'clip_values_det_limits': False,
        })
        self.exposure = exposure.Exposure(self.detector, filters.F140W(), self.planet, self.exp_info)
        self.exposure.add_read(self.detector.gen_pixel_array(self.SUBARRAY, light_sensitive=False))
        S = self.SUBARRAY
        ax = np.arange(S, dtype=float) + 0.5
        x, y = np.meshgrid(ax, ax)
        x0 = x_ref - (507.0 - S / 2.0)
        y0 = y_ref - (507.0 - S / 2.0)
        sigma = 2.0
        image = 10000.0 * np.exp(-((x0 - x) ** 2 + (y0 - y) ** 2) / (2.0 * sigma * sigma))
        self.exposure.add_read(image, {'read_exp_time': 0 * u.s, 'cumulative_exp_time': 0 * u.s,
                                       'CRPIX1': -5})
        return self.exposure

    def staring_frame(self, x_ref, y_ref, x_jitter, y_jitter, wl, stellar_flux, planet_signal,
                      sample_mid_points, sample_durations, read_index, noise_mean, noise_std,
                      add_dark, add_flat, cosmic_rate, sky_background, scale_factor,
                      add_gain_variations, add_non_linear, clip_values_det_limits, add_read_noise,
                      add_stellar_noise, add_initial_bias, progress_bar, threads=2):
        """A stationary scan: speed 0 and a sample rate so long that every read
        interval is one sub-sample (exposure_generator.py:146-176)."""
        self.exposure = self.scanning_frame(
            x_ref, y_ref, x_jitter, y_jitter, wl, stellar_flux, planet_signal,
            0 * u.pixel / u.s, 1 * u.year, sample_mid_points, sample_durations, read_index, None,
            noise_mean, noise_std, add_dark, add_flat, cosmic_rate, sky_background, scale_factor,
            add_gain_variations, add_non_linear, clip_values_det_limits, add_read_noise,
            add_stellar_noise, add_initial_bias, progress_bar, threads)
        return self.exposure

    # ------------------------------------------------------------------
    def _gen_scanning_sample_times(self, sample_rate):
        """(starts, mid points, durations) [ms Quantities] and the index of the
        last sub-sample of every read: sampling restarts after each read and a
        read's final sub-sample is cut to the remainder (:531-579)."""
        rate = float(u.value_in(sample_rate, u.ms))
        read_times = np.asarray(u.value_in(self.read_times, u.ms), dtype=float)
        pieces, read_index, last, prev = [], [], -1, 0.
        for t in read_times:
            starts = np.arange(prev, t, rate)
            pieces.append(starts)
            last += len(starts)
            read_index.append(last)
            prev = t
        starts = np.concatenate(pieces)
        ends = np.roll(starts, -1)
        ends[-1] = read_times[-1]
        durations = ends - starts
        mids = starts + (durations / 2)
        return starts * u.ms, mids * u.ms, durations * u.ms, read_index

    def _gen_sample_yref(self, y_ref, mid_points, scan_speed):
        return y_ref + _ms(mid_points) * float(u.value_in(scan_speed, u.pixel / u.ms))

    def _gen_zero_read(self, add_initial_bias=True):
        zero = self.detector.gen_pixel_array(self.SUBARRAY, light_sensitive=False)
        if self.SUBARRAY == 256 and add_initial_bias:
            zero = zero + self.detector.get_initial_bias()
        return zero, {'cumulative_exp_time': 0 * u.s, 'read_exp_time': 0 * u.s, 'CRPIX1': 0}

    @property
    def photons(self):
        """Electrons thrown by the last exposure (incl. those that left the frame)."""
        return self._run.photons()

    # ------------------------------------------------------------------
    def _rng_mode(self):
        mode = self.rng if self.rng is not None else params.rng
        if mode not in ('philox', 'numpy'):
            raise ValueError("rng must be 'philox' or 'numpy', got {!r}".format(mode))
        return mode

    def _default_key(self, compat=False):
        seed = params.seed
        if seed is None:
            # compat mode must not consume the numpy stream (its order is the
            # reference's); native mode takes one draw so np.random.seed() still
            # makes a run reproducible
            seed = 0 if compat else int(np.random.randint(0, 2 ** 31 - 1))
        return (int(seed) & 0xffffffff, zlib.crc32(str(self.exp_info['filename']).encode()) & 0xffffffff)

    def _device_planes(self, eng, add_gain_variations, sky_background, add_non_linear, zero_read):
        """Resident float64 device planes in the bordered F x F layout."""
        S, det, g = self.SUBARRAY, self.detector, self.grism
        L, F = det.light_side(S), det.full_side(S)
        sky = gain = nl = None
        if sky_background:
            sky = eng.cached_plane(('sky', g.name, g.sky_file_name, S),
                                   lambda: eng.to_dev(eng.bordered(g.get_master_sky(L), F)))
        if add_gain_variations:
            gain = eng.cached_plane(('gain', det.gain_file_name, S),
                                    lambda: eng.to_dev(eng.bordered(det.get_gain(S), F, fill=1.0)))
        if add_non_linear:
            nl = eng.cached_plane(('nl', det.non_linear_file_name, F),
                                  lambda: tuple(eng.to_dev(p) for p in det.non_linear_planes(F)))
        zero = None
        if zero_read is not None:
            zero = eng.cached_plane(('bias', self.SUBARRAY), lambda: eng.to_dev(zero_read))
        return sky, gain, nl, zero

    def _dark_stack(self, R):
        """(dark[R][F][F], err[R][F][F]) host float64 stacks: read r uses the
        super-dark extension of NSAMP = r + 1 (exposure.py:70-80)."""
        det = self.detector
        planes = [det.dark_planes(r + 1, self.SUBARRAY, self.SAMPSEQ) for r in range(1, R + 1)]
        return (np.stack([np.asarray(p[0], dtype=np.float64) for p in planes]),
                np.stack([np.asarray(p[1], dtype=np.float64) for p in planes]))

    # ------------------------------------------------------------------
    def scanning_frame(self, x_ref, y_ref, x_jitter, y_jitter, wl, stellar_flux, planet_signal,
                       scan_speed, sample_rate, sample_mid_points=None, sample_durations=None,
                       read_index=None, ssv_generator=None, noise_mean=False, noise_std=False,
                       add_dark=True, add_flat=True, cosmic_rate=None,
                       sky_background=1 * u.count / u.s, scale_factor=None,
                       add_gain_variations=True, add_non_linear=True, clip_values_det_limits=True,
                       add_read_noise=True, add_stellar_noise=True, add_initial_bias=True,
                       progress_bar=None, threads=2, rng_key=None, exact_newton=None,
                       out_dtype=np.float64, device_result=False, electron_normals=None):
        """Generate a spatially scanned exposure (see the module docstring).

        Units of bare numbers: ``wl`` micron, ``stellar_flux`` erg/(angstrom s
        cm^2), ``scan_speed`` pixel/s, ``sample_rate`` and sample times ms,
        ``sky_background`` count/s.  Extra keywords over the reference:
        ``rng_key`` (Philox key pair), ``exact_newton`` (reference's global
        Newton stopping rule; default on in 'numpy' mode), ``out_dtype``
        (float64 like the reference, or float32), ``device_result`` (leave the
        reads in HBM as ``exposure.device_reads`` [NSAMP][F][F] and skip the
        device->host copy; ``wl`` / ``stellar_flux`` stay host arrays but
        ``planet_signal`` may be a CUDA tensor already resident in HBM),
        ``electron_normals`` ('numpy' mode only: callable(sample index, n electrons)
        -> the table A[2n] of PSF() -- x normals then y normals, pyparallel_menu.c:
        55-62 -- used instead of the rand_r stream: the fully deterministic photon
        list of BASELINE configs[1])."""
        from . import _lib
        from .engine import DeviceEngine, ExposureRun

        start_time = time.time()
        mode = self._rng_mode()
        compat = mode == 'numpy'
        if exact_newton is None:
            exact_newton = compat
        self.transmission_spectroscopy = planet_signal is not None

        if not u.is_quantity(scan_speed):
            scan_speed = scan_speed * (u.pixel / u.s)
        scan_speed_ms = float(scan_speed.to(u.pixel / u.ms).value)
        sample_rate_q = sample_rate if u.is_quantity(sample_rate) else sample_rate * u.ms
        sample_rate_q = sample_rate_q.to(u.ms)

        if sample_mid_points is None and sample_durations is None and read_index is None:
            _, sample_mid_points, sample_durations, read_index = \
                self._gen_scanning_sample_times(sample_rate_q)

        mid_ms = _ms(sample_mid_points)
        s_y_refs = y_ref + mid_ms * scan_speed_ms

        key = tuple(int(k) for k in (rng_key if rng_key is not None else self._default_key(compat)))
        if ssv_generator is not None:
            # the generators draw from numpy's global stream (scan_speed_varations.py:45, 100-132).
            # compat: that IS the reference's stream and its order (A.7).  native: the draws of an
            # exposure must depend on its key only -- not on how many exposures this process has
            # generated before, i.e. not on how a visit is spread over GPUs -- so the global
            # stream is keyed for the call and put back afterwards.
            saved = None
            if not compat:
                saved = np.random.get_state()
                np.random.seed([key[0], key[1], 0x55F])
            try:
                if isinstance(ssv_generator, scan_speed_varations.SSVModulatedSine):
                    sample_durations, read_index = ssv_generator.get_subsample_exposure_times(
                        s_y_refs, sample_durations, self.read_times, sample_rate_q)
                else:
                    sample_durations = ssv_generator.get_subsample_exposure_times(
                        s_y_refs, sample_durations, self.read_times, sample_rate_q)
            finally:
                if saved is not None:
                    np.random.set_state(saved)
        dur_ms = _ms(sample_durations)

        self.exp_info.update({
            'SCAN': True, 'SCAN_DIR': 1, 'samp_rate': sample_rate_q, 'x_ref': x_ref, 'y_ref': y_ref,
            'noise_mean': noise_mean, 'noise_std': noise_std, 'add_dark': add_dark,
            'add_flat': add_flat, 'add_gain': add_gain_variations, 'add_non_linear': add_non_linear,
            'add_stellar_noise': add_stellar_noise, 'cosmic_rate': cosmic_rate,
            'sky_background': sky_background, 'scale_factor': scale_factor,
            'clip_values_det_limits': clip_values_det_limits, 'rng': mode,
        })
        self.exposure = exposure.Exposure(self.detector, self.grism, self.planet, self.exp_info)
        # the zero read is the initial bias for SUBARRAY 256 and zeros otherwise
        # (:446-466); the all-zero case never materialises an F x F host array
        has_bias = bool(self.SUBARRAY == 256 and add_initial_bias)
        zero_read = self.detector.get_initial_bias() if has_bias else None
        zero_read_info = {'cumulative_exp_time': 0 * u.s, 'read_exp_time': 0 * u.s, 'CRPIX1': 0}

        if progress_bar is not None:
            progress_bar.print_status_line(progress_bar.progress_line + ' (device)')

        det = self.detector
        S = self.SUBARRAY
        L, F = det.light_side(S), det.full_side(S)
        num_samples = len(mid_ms)
        read_index = [int(i) for i in read_index]
        R = len(read_index)
        read_times_s = np.asarray(u.value_in(self.read_times, u.s), dtype=np.float64)
        if R != len(read_times_s):
            raise ValueError("read_index must have one entry per non-zero read")
        if len(dur_ms) < num_samples:      # bad SSV output: missing durations count as 0 (:340-342)
            dur_ms = np.concatenate([dur_ms, np.zeros(num_samples - len(dur_ms))])
        dt_s = np.diff(np.concatenate([[0.0], read_times_s]))

        eng = DeviceEngine.get(self.device)
        eng.admit()

        # ---- per-sub-sample seeds and pointing jitter (:327-329) ---------------
        if compat:
            s_rand_seeds = np.random.randint(0, 100000, num_samples)
            s_x_jitter = np.random.normal(0, x_jitter, num_samples)
            s_y_jitter = np.random.normal(0, y_jitter, num_samples)
        else:
            g = np.random.Generator(np.random.Philox(key=(key[0] << 32) | key[1]))
            s_rand_seeds = None
            s_x_jitter = g.normal(0, 1, num_samples) * x_jitter
            s_y_jitter = g.normal(0, 1, num_samples) * y_jitter

        # ---- spectrum crop (:332-334) and stage 1 --------------------------------
        wl_um = np.asarray(u.value_in(wl, u.micron), dtype=np.float64)
        lo = float(u.value_in(self.grism.wl_limits[0], u.micron))
        hi = float(u.value_in(self.grism.wl_limits[-1], u.micron))
        i0, i1 = tools.crop_spectrum_ind(lo, hi, wl_um)
        flux = np.asarray(getattr(stellar_flux, 'value', stellar_flux), dtype=np.float64)[i0:i1]
        depth = None
        if planet_signal is not None:
            if hasattr(planet_signal, 'is_cuda') or hasattr(planet_signal, 'coef') or hasattr(planet_signal, 'row'):
                depth = planet_signal          # CUDA tensor, lightcurve.ChebyshevSignal / SeparableSignal
            else:
                depth = np.asarray(planet_signal)
            if depth.ndim != 2 or depth.shape[0] < num_samples:
                raise ValueError("planet_signal must be [n_samples][n_wl]")
            depth = depth[:num_samples]
        aux = {'dt': dt_s}
        native_cosmics = None
        host_cosmics = None
        if not compat and cosmic_rate is not None:
            # hit list of the whole exposure (cosmic_rays.py:88-139 per read interval),
            # drawn up front so it rides in the exposure's first small upload
            g = np.random.Generator(np.random.Philox(key=((key[0] << 32) | key[1]) ^ 0xC05B1C))
            rd, rows, cols, en = [], [], [], []
            for r in range(R):
                n = g.poisson(cosmic_rate / (1024. * 1024.) * (L * L) * dt_s[r])
                rd.append(np.full(n, r, np.int32))
                en.append(g.integers(10000, 35000, n).astype(np.float64))
                rows.append(g.integers(0, L, n))
                cols.append(g.integers(0, L, n))
            rows, cols = np.concatenate(rows), np.concatenate(cols)
            if len(rows):
                host_cosmics = (((rows + BORDER) * F + (cols + BORDER)).astype(np.int32), np.concatenate(rd),
                                np.concatenate(en))
                aux['cos_pix'], aux['cos_rd'], aux['cos_en'] = host_cosmics
                native_cosmics = True

        sky_rate = float(u.value_in(sky_background, u.count / u.s)) if sky_background else 0.0
        use_noise = bool(noise_mean and noise_std)
        if not compat and params.use_context and params.direct_accumulation and not exact_newton:
            return self._scan_through_context(
                eng, start_time, wl_um[i0:i1], flux, depth, i0, x_ref + s_x_jitter, s_y_refs + s_y_jitter,
                dur_ms, dt_s, read_index, scale_factor, key, add_stellar_noise, add_flat, sky_rate,
                add_gain_variations, add_dark, add_non_linear, clip_values_det_limits, add_read_noise,
                zero_read, (noise_mean, noise_std) if use_noise else (0.0, 0.0), host_cosmics, out_dtype,
                device_result, read_times_s, zero_read_info, progress_bar)
        if hasattr(depth, 'row'):
            depth = depth.to_array()           # the stage-by-stage path takes the dense form
        run = ExposureRun(eng, self.grism, S, wl_um[i0:i1], flux, depth, i0,
                          x_ref + s_x_jitter, s_y_refs + s_y_jitter, dur_ms, scale_factor,
                          np.asarray(read_index, dtype=np.int32), aux=aux)
        self._run = run

        draws = {}
        cosmics = None
        dark = None          # host (dark, err) stacks: only the compat draws need them
        d_dark = None        # resident device stacks (native mode)
        if add_dark:
            try:
                if compat:
                    dark = self._dark_stack(R)
                else:
                    def upload():
                        host = self._dark_stack(R)
                        return eng.to_dev(host[0]), eng.to_dev(host[1])
                    d_dark = eng.cached_plane(('dark', S, self.SAMPSEQ, R), upload)
            except WFC3SimNoDarkFileError:
                warnings.warn("No Dark file found for SAMPSEQ = {}, SUBARRAY={} - Switching Dark "
                              "Off".format(self.SAMPSEQ, S), WFC3SimNoDarkFileWarning)
                self.exposure.exp_info['add_dark'] = False
                add_dark = False

        if compat:
            # the reference's numpy stream, in its order
            expected = run.expected_host()
            counts = np.empty((num_samples, run.W), dtype=np.int32)
            r = 0
            noise_d, sky_d, hits = [], [], []
            ends = set(read_index)
            for i in range(num_samples):
                if add_stellar_noise:
                    counts[i] = np.random.poisson(expected[i])
                else:
                    counts[i] = np.round(expected[i])
                if i in ends:
                    dt = dt_s[r]
                    if use_noise:
                        noise_d.append(np.random.normal(noise_mean * dt, noise_std * dt, (L, L)))
                    if sky_rate:
                        sky = np.asarray(self.grism.get_master_sky(L), dtype=np.float32)
                        lam = sky * np.float32(sky_rate * dt)      # in-place float32 product (:493)
                        sky_d.append(np.random.poisson(lam))
                    if cosmic_rate is not None:
                        gen = cosmic_rays.MinMaxPossionCosmicGenerator(cosmic_rate)
                        rows, cols, en = gen.cosmic_hits(dt, L)
                        hits.append((np.full(len(en), r, np.int32), rows, cols, en))
                    r += 1
            run.counts(_lib.COUNT_NONE, counts=counts)
            if use_noise:
                draws['noise'] = np.stack(noise_d)
            if sky_rate:
                draws['sky'] = np.stack(sky_d).astype(np.float64)
            if hits:
                rd = np.concatenate([h[0] for h in hits])
                rows = np.concatenate([h[1] for h in hits])
                cols = np.concatenate([h[2] for h in hits])
                en = np.concatenate([h[3] for h in hits])
                cosmics = ((rows + BORDER) * F + (cols + BORDER), rd, en)
            if add_dark:
                draws['dark'] = np.stack([np.random.normal(dark[0][k], dark[1][k]) for k in range(R)])
                dark = None
            if add_read_noise:
                draws['rn'] = np.random.standard_normal((R + 1, F, F))
            if electron_normals is not None:
                totals = counts.astype(np.int64).sum(axis=1)
                tables = [np.asarray(electron_normals(i, int(totals[i])), dtype=np.float64)
                          for i in range(num_samples)]
                for i, t in enumerate(tables):
                    if t.size != 2 * int(totals[i]):
                        raise ValueError("electron_normals({}, n) must return 2*n values".format(i))
                run.throw(_lib.RNG_HOST, normals=np.concatenate(tables) if tables else np.zeros(0),
                          add_flat=add_flat)
            else:
                run.throw(_lib.RNG_RANDR, seeds=s_rand_seeds, threads=threads, add_flat=add_flat)
        else:
            run.counts(_lib.COUNT_POISSON if add_stellar_noise else _lib.COUNT_ROUND, key=key)
            if native_cosmics:
                cosmics = (run.aux['cos_pix'], run.aux['cos_rd'], run.aux['cos_en'])
            if params.direct_accumulation:
                run.throw_direct(key=key, add_flat=add_flat)
            else:
                run.throw(_lib.RNG_PHILOX, key=key, add_flat=add_flat)

        sky_p, gain_p, nl_p, zero_p = self._device_planes(
            eng, add_gain_variations, sky_rate, add_non_linear, zero_read)
        out = run.reads(
            run.aux['dt'], key=key, sky_rate=sky_rate, sky_plane=sky_p, gain_plane=gain_p, zero=zero_p,
            dark=d_dark, nl_planes=nl_p, noise=(noise_mean, noise_std) if use_noise else (0.0, 0.0),
            clip=(det.min_counts, det.max_counts) if clip_values_det_limits else None,
            read_noise=det.read_noise if add_read_noise else 0.0, cosmics=cosmics, draws=draws,
            exact_newton=bool(exact_newton and add_non_linear),
            out_f32=(np.dtype(out_dtype) == np.float32), const_gain=det.constant_gain,
            fast_math=not compat)

        return self._hand_over(eng, out, run.lost, device_result, start_time, read_times_s, zero_read_info,
                               num_samples, progress_bar)

    def _scan_through_context(self, eng, start_time, wl_um, flux, depth, depth_col0, xr, yr, dur_ms, dt_s,
                              read_index, scale_factor, key, add_stellar_noise, add_flat, sky_rate,
                              add_gain_variations, add_dark, add_non_linear, clip, add_read_noise, zero_read,
                              noise, cosmics, out_dtype, device_result, read_times_s, zero_read_info,
                              progress_bar):
        """Native mode: the whole exposure is ONE call on the configuration's resident context
        (include/wayne_b200.h, wb200_exposure_run) -- stage 1, counts, thrower + flat, ramp pass."""
        from . import _lib
        det, S = self.detector, self.SUBARRAY
        R = len(read_index)
        ctx = eng.exposure_context(self.grism, det, S, self.SAMPSEQ)
        if add_dark:
            try:
                ctx.ensure_planes(False, False, False, False, R, None)
            except WFC3SimNoDarkFileError:
                warnings.warn("No Dark file found for SAMPSEQ = {}, SUBARRAY={} - Switching Dark "
                              "Off".format(self.SAMPSEQ, S), WFC3SimNoDarkFileWarning)
                self.exposure.exp_info['add_dark'] = False
                add_dark = False
        ctx.ensure_planes(add_flat, bool(sky_rate), add_gain_variations, add_non_linear, 0, zero_read)
        out, run = ctx.run(wl_um, flux, depth, depth_col0, xr, yr, dur_ms, dt_s,
                           np.asarray(read_index, dtype=np.int32), scale_factor, key,
                           _lib.COUNT_POISSON if add_stellar_noise else _lib.COUNT_ROUND,
                           add_flat, sky_rate, add_gain_variations, add_dark, add_non_linear, clip,
                           add_read_noise, zero_read is not None, noise, cosmics,
                           np.dtype(out_dtype) == np.float32)
        self._run = run
        return self._hand_over(eng, out, run.stats, device_result, start_time, read_times_s, zero_read_info,
                               len(xr), progress_bar, stats_of=run)

    def _hand_over(self, eng, out, lost, device_result, start_time, read_times_s, zero_read_info,
                   num_samples, progress_bar, stats_of=None):
        """Leave the reads in HBM (device_result) or queue the single device->host copy into
        pooled pinned memory; exposure.reads waits for it on first access."""
        from . import _lib
        R = len(read_times_s)
        if device_result:
            eng.retire()
            self.exposure.device_reads = out
            self.exp_info['sim_time'] = (time.time() - start_time) * u.s
            return self.exposure
        done, reads_host, lost = eng.fetch_async(out, small=lost)
        nsamp = self.NSAMP
        if stats_of is not None:
            # the context path's small tensor is the exposure's statistics (thrown / binned /
            # dropped electrons), not a lost-electron counter: it rides the same download
            stats_of.stats_host, stats_of.stats_done = lost, done
            lost = None

        def materialize(exp):
            done.synchronize()
            n_lost = int(lost[0]) if lost is not None else 0
            if n_lost:
                raise _lib.WayneB200Error(
                    "{} electrons fell outside their sub-sample window".format(n_lost))
            exp.add_read(reads_host[0], zero_read_info)
            prev = 0.0
            for r in range(R):
                exp.add_read(reads_host[r + 1], {
                    'cumulative_exp_time': read_times_s[r] * u.s,
                    'read_exp_time': (read_times_s[r] - prev) * u.s,
                    'CRPIX1': 0,
                })
                prev = read_times_s[r]
            assert len(exp._reads) == nsamp
            exp.exp_info['sim_time'] = (time.time() - start_time) * u.s

        self.exposure._pending = materialize

        if progress_bar is not None:
            progress_bar.print_status_line(
                progress_bar.progress_line + ' (samp {0}/{0})'.format(num_samples))
        self.exp_info['sim_time'] = (time.time() - start_time) * u.s
        return self.exposure
