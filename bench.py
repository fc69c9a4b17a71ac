#!/usr/bin/env python
"""bench.py -- exposures/s and photons/s of the exposure-synthesis path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--workload c4|c1|tiny] [--no-cpu] [--no-extras]

A "step" is one full exposure (stages 1-4: trace/dispersion/sensitivity, Philox
Poisson counts, electron throw + binning, flat-field gather, fused per-pixel
ramp pass).  Workload at N=1 = BASELINE.json configs[3], the configuration the
metric is quoted on: G141 spatial scan, 1024^2 full frame, NSAMP=15 RAPID,
10 ms sub-samples (4116 of them), 4096 wavelength bins, ~1e9 photons.

  value   exposures/s with the exposure's inputs already resident in HBM and the
          reads left in HBM (CUDA events on the launching stream, max over ranks)
  e2e     the same through ExposureGenerator.scanning_frame with HOST buffers:
          pinned host->device copy of the inputs and device->host copy of the
          NSAMP reads inside the timed region (copies of consecutive exposures
          overlap the kernels: upload / compute / download streams; every
          exposure's reads are touched on the host before the clock stops)
  roofline / roofline_hbm   per-kernel, durations from CUDA events recorded around each
          launch, one kernel at a time, over 4-12 exposures of the same inputs right AFTER the
          timed region (per-stage events keep an exposure on one stream; inside the timed
          region consecutive exposures may overlap) -- `stage_ms`
  e2e.copy_ceiling / frac_of_copy_ceiling   what THIS box moves with concurrent pinned copies
          of the same sizes on the same GPUs and nothing else running (tools/copy_ceiling.py,
          0.6 s per mode, all ranks at once): the bound of every e2e figure, and the e2e
          figure as a fraction of it
  cpu_baseline   the oracle (unmodified reference C kernel when it was compiled
          in the build container, else the C port) + numpy restatement, timed on
          this host on a bounded sample of the same exposure

  e2e.driver_form / e2e.separable_form   the same exposure with the planet signal handed
          over the way the visit driver does (Chebyshev coefficients) or as its two
          factors (lightcurve.SeparableSignal): no 135 MB upload
  multi_visit   BASELINE configs[4]: 8 visits x 128 exposures of this shape, pointing random
          walk sigma = 0.02 px, partitioned exposure-wise over the ranks through
          sharding.run_sharded (strong scaling: the batch is fixed), summaries gathered
  psf_dropin    the PSF() / apply_psf() drop-in at one sub-sample's shape next to the
          reference's own C kernel and Cython wrapper (rank 0, N = 1)

N > 1 (torchrun): exposures are independent, so every rank runs its own
exposures (weak scaling, no data-path collective); value = all ranks' exposures
/ max-over-ranks time.

--impl reference: the reference's CPU path (oracle/) alone, rank 0 only: the first timed
step is one FULL exposure (every sub-sample), the others bounded samples of it.
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: grism, SUBARRAY, NSAMP, SAMPSEQ, scan px/s, sample ms, W, photons, x_ref, y_ref
    'c4': dict(grism='G141', sub=1024, nsamp=15, seq='RAPID', scan=20.0, rate=10.0, W=4096,
               photons=1.0e9, x_ref=330.0, y_ref=110.0,
               desc='configs[3]: G141 spatial scan 1024x1024 full frame, NSAMP=15 RAPID (41.05 s), '
                    '20 px/s, 10 ms sub-samples, 4096 wavelength bins, ~1e9 photons/exposure, '
                    'flat+sky+cosmics+gain+dark+non-linearity+clip+read noise on, SSVSine, visit trend'),
    'c1': dict(grism='G141', sub=256, nsamp=5, seq='SPARS10', scan=7.4325, rate=10.0, W=4494,
               photons=3.0e8, x_ref=404.497, y_ref=457.427,
               desc='configs[0]-shaped: G141 scan 256 subarray NSAMP=5 SPARS10, 2233 sub-samples x 4494 bins'),
    'tiny': dict(grism='G141', sub=256, nsamp=5, seq='SPARS10', scan=7.4325, rate=200.0, W=512,
                 photons=2.0e6, x_ref=404.497, y_ref=457.427, desc='smoke-sized scan'),
}


def bench_config(wk, n_sub, n_bins):
    """The `config` object of the JSON line: identical keys and values for both arms."""
    return {'workload': wk['desc'], 'grism': wk['grism'], 'subarray': wk['sub'], 'nsamp': wk['nsamp'],
            'sampseq': wk['seq'], 'n_subsamples': int(n_sub), 'n_bins': int(n_bins),
            'photons_per_exposure_target': wk['photons'], 'out_dtype': 'float64',
            'terms': 'flat+sky+cosmics+gain+dark+non-linearity+clip+read noise, SSVSine, visit trend',
            'l2': 'no flush: the per-exposure working set (interval planes 117 MB + dark 117 MB + reads 126 MB '
                  '[+ planet signal 135 MB]) exceeds the 126 MB L2'}


def calibration_dir(wk):
    from wayne_b200 import calibration, params
    d = os.environ.get('WAYNE_CALB_DIR') or os.path.join(tempfile.gettempdir(), 'wayne_b200_synth_calb')
    rank = int(os.environ.get('LOCAL_RANK', '0'))
    if rank == 0:
        calibration.write_synthetic_calibration(d, modes=((wk['sub'], wk['seq']),))
        open(os.path.join(d, '.ready_%s_%s' % (wk['sub'], wk['seq'])), 'w').close()
    else:
        while not os.path.exists(os.path.join(d, '.ready_%s_%s' % (wk['sub'], wk['seq']))):
            time.sleep(0.2)
    params.set_calibration_dir(d)
    return d


def make_inputs(wk, exposure_index=0):
    """Host inputs of one exposure of the workload (seeded, synthetic)."""
    from wayne import detector, grism
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    g = grism.G141() if wk['grism'] == 'G141' else grism.G102()
    det = detector.WFC3_IR()
    eg = ExposureGenerator(det, g, wk['nsamp'], wk['seq'], wk['sub'], None)
    _, mid, dur, ri = eg._gen_scanning_sample_times(wk['rate'] * u.ms)
    lo, hi = (float(u.value_in(v, u.micron)) for v in g.wl_limits)
    wl = np.exp(np.linspace(np.log(lo), np.log(hi), wk['W']))       # log-spaced, all inside the limits
    rng = np.random.default_rng(20170410)
    x = 1.4388e4 / (wl * 6065.0)
    bb = 1.0 / (wl ** 5 * (np.exp(x) - 1.0))
    flux = bb / bb.max() * (1 + 0.01 * rng.standard_normal(len(wl)))
    # scale to the target photon count: sum over (sample, bin) of expected electrons
    sens_wl, sens_val = g._load_sens()
    from wayne import tools
    per_ms = flux * np.interp(wl, sens_wl, sens_val) * tools.bin_centers_to_widths(wl) * 1e4 * 1e-3
    total = per_ms.sum() * float(np.sum(u.value_in(dur, u.ms)))
    flux = flux * (wk['photons'] / total)
    depth0 = 0.0146 * (1 + 0.01 * np.sin(12 * wl))
    lc = 0.5 * (1 + np.tanh((np.linspace(-1, 1, len(mid)) + 0.2 * exposure_index) * 3))
    return dict(eg_args=(det, g, wk['nsamp'], wk['seq'], wk['sub'], None), wl=wl, flux=flux,
                depth0=depth0, lightcurve=lc, mid=mid, dur=dur, read_index=ri,
                read_times=np.asarray(u.value_in(eg.read_times, u.s), dtype=float))


def frame_kwargs(wk, i):
    from wayne import units as u
    from wayne.trend_generators.scan_speed_varations import SSVSine
    rw = np.random.default_rng(1963 + i)
    return dict(x_ref=wk['x_ref'] + 0.02 * rw.standard_normal(), y_ref=wk['y_ref'] + 0.02 * rw.standard_normal(),
                x_jitter=0.025, y_jitter=0.025, scan_speed=wk['scan'] * u.pixel / u.s,
                sample_rate=wk['rate'] * u.ms, ssv_generator=SSVSine(1.5, 1.1, 0), cosmic_rate=11.,
                sky_background=5.5 * u.count / u.s, scale_factor=1.0 - 1e-4 * (i % 100))   # (i carries rank * 100000)


class ClockSampler(object):
    """SM clock and throttle reasons sampled DURING the timed region.

    In-process NVML (nvidia_ml_py) from a thread: spawning `nvidia-smi -lms` next
    to the bench stalls the CUDA driver for 50-100 ms at a time on this box (seen
    as host-issue outliers), a lightweight NVML query every 25 ms does not."""

    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20),
               ('sw_power_cap', 0x4))

    def __init__(self, index, period=0.025):
        import threading
        self.samples, self.mask, self.max_mhz, self.err = [], 0, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as exc:          # noqa: BLE001
            self.err = 'nvml unavailable: %s' % exc
            return

        def loop():
            while not self._stop.is_set():
                try:
                    self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                except Exception:         # noqa: BLE001
                    pass
                self._stop.wait(period)

        self.t = threading.Thread(target=loop, daemon=True)
        self.t.start()

    def reset(self):
        """Start of the timed region: forget the samples taken so far."""
        self.samples, self.mask = [], 0

    def stop(self):
        if self.err:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [self.err]}
        self._stop.set()
        self.t.join()
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'samples': len(self.samples),
                'reasons': [n for n, bit in self.REASONS if self.mask & bit],
                'source': 'NVML, in-process, every 25 ms during the timed region'}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------
# CPU reference arm (oracle/): bounded sample of the same exposure
# ---------------------------------------------------------------------------
def psf_case_of(wk, inp, electrons=None, seed=3):
    """PSF() inputs of ONE sub-sample of the workload (oracle-side arithmetic only):
    counts, x, y, ratio, sigma_l, sigma_h, frame side."""
    from oracle import exposure_oracle as E
    from tests import harness
    from wayne import units as u
    cal = harness.oracle_calibration(wk['grism'], dark_mode=None)
    a_, b_ = (E.G141_TRACE, E.G141_WLSOL) if wk['grism'] == 'G141' else (E.G102_TRACE, E.G102_WLSOL)
    lim = E.WL_LIMITS[wk['grism']]
    i0, i1 = E.crop_spectrum_ind(lim[0], lim[1], inp['wl'])
    s_wl = inp['wl'][i0:i1]
    ratio, sigl, sigh, sens, dwl = E.bin_tables(s_wl, cal['sens_wl_um'], cal['sens_val'])
    L_ = 1014 if wk['sub'] == 1024 else wk['sub']
    tr = E.Trace(wk['x_ref'], wk['y_ref'] + 50.0, a_, b_)
    sub_scale = 507 - wk['sub'] // 2
    xs, ys = tr.wl_to_x(s_wl) - sub_scale, tr.wl_to_y(s_wl) - sub_scale
    dur = np.asarray(u.value_in(inp['dur'], u.ms))
    mean = E.expected_counts(inp['flux'][i0:i1], None, sens, dwl, float(np.median(dur)), None)
    if electrons is not None:
        mean = mean * (electrons / mean.sum())
    cnt = np.random.RandomState(seed).poisson(mean).astype(np.int32)
    return cnt, xs, ys, ratio, sigl, sigh, L_


def time_reference_psf(case, ncpu):
    """The reference's native kernel alone on one sub-sample's inputs: PSF() through ctypes per
    OpenMP team size, and apply_psf() through the unmodified Cython wrapper (what the reference
    pays per sub-sample, pyparallel.pyx:14-38).  Seconds per call."""
    from oracle import psf as OP
    cnt, xs, ys, ratio, sigl, sigh, L_ = case
    out = {'electrons_per_call': int(cnt.sum()), 'frame': int(L_), 'bins': int(len(cnt)), 'psf_s_by_threads': {}}
    for t in [t for t in (1, 2, 4, 8, 16, 32) if t <= ncpu]:
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            OP.psf_reference(cnt, xs, ys, ratio, sigl, sigh, L_, L_, 7, t)
            el = time.perf_counter() - t0
            best = el if best is None or el < best else best
        out['psf_s_by_threads'][str(t)] = best
    pyx = OP.reference_pyparallel()
    if pyx is not None:
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            pyx.apply_psf(cnt, xs, ys, ratio, sigl, sigh, L_, L_, 7, 1)
            el = time.perf_counter() - t0
            best = el if best is None or el < best else best
        out['apply_psf_cython_s_1thread'] = best
    return out


def cpu_reference_exposure(wk, inp, n_sub, threads=None, psf_alone=True):
    """Time the reference's CPU path on ``n_sub`` evenly spread sub-samples (None: ALL of them,
    a full un-sampled exposure) plus ALL read reductions and the post-exposure chain; a sampled
    run extrapolates the sub-sample part to the exposure's N sub-samples.
    Returns (seconds per exposure, info)."""
    from oracle import exposure_oracle as E
    from oracle import psf as OP
    from tests import harness
    kind = 'reference' if OP.have_reference() else 'port'
    cal = harness.oracle_calibration(wk['grism'], dark_mode=(wk['sub'], wk['seq']), nsamp=wk['nsamp'])
    from wayne import units as u
    mid = np.asarray(u.value_in(inp['mid'], u.ms))
    dur = np.asarray(u.value_in(inp['dur'], u.ms))
    ri = inp['read_index']
    N = len(mid)
    if n_sub is None or n_sub >= N:
        picks, new_ri = np.arange(N), list(ri)
    else:
        # evenly spread sub-samples, at least one per read, each read closed by its last pick
        picks, new_ri, first = [], [], 0
        per_read = max(1, n_sub // len(ri))
        for last in ri:
            idx = np.unique(np.linspace(first, last, per_read).round().astype(int))
            picks.extend(idx.tolist())
            new_ri.append(len(picks) - 1)
            first = last + 1
        picks = np.array(picks)
    depth = inp['depth0'][None, :] * inp['lightcurve'][picks][:, None]
    ncpu = os.cpu_count() or 1
    if threads is None:
        if kind == 'reference':
            # the reference's OpenMP team only speeds up the normal table: pick
            # the best team size on one sub-sample
            best, threads = None, 1
            for t in [t for t in (1, 2, 4, 8, 16) if t <= ncpu]:
                t0 = time.perf_counter()
                E.scanning_frame(cal, wk['grism'], wk['sub'], inp['read_times'][:1], inp['wl'], inp['flux'],
                                 None, wk['x_ref'], wk['y_ref'], 0.0, 0.0, 0.0, wk['rate'],
                                 np.random.RandomState(1), add_dark=False, sky_background=0,
                                 add_non_linear=False, add_read_noise=False, threads=t, psf=kind,
                                 sample_times=(mid[:2], dur[:2], [1]))
                el = time.perf_counter() - t0
                if best is None or el < best:
                    best, threads = el, t
        else:
            threads = 1
    psf_info = None
    if kind == 'reference' and psf_alone:
        psf_info = time_reference_psf(psf_case_of(wk, inp), ncpu)
    kw = frame_kwargs(wk, 0)
    t0 = time.perf_counter()
    o = E.scanning_frame(cal, wk['grism'], wk['sub'], inp['read_times'], inp['wl'], inp['flux'], depth,
                         kw['x_ref'], kw['y_ref'], 0.025, 0.025, wk['scan'] * 0.001, wk['rate'],
                         np.random.RandomState(1963), ssv=(1.5, 1.1, 0), cosmic_rate=11.,
                         sky_background=5.5, scale_factor=kw['scale_factor'], threads=threads, psf=kind,
                         sample_times=(mid[picks], dur[picks], new_ri))
    wall = time.perf_counter() - t0
    tm = o['timing']
    scale = N / float(len(picks))
    per_exposure = tm['subsamples'] * scale + tm['reads'] + tm['post']
    photons = o['photons'] * scale
    full = len(picks) == N
    info = {'kind': kind, 'cores': int(threads), 'full': full,
            'sample': ('all %d sub-samples (a full, un-sampled exposure) + all %d read reductions + '
                       'post-exposure chain; %.1f s of CPU work' % (N, len(ri), wall)) if full else
                      ('%d of %d sub-samples (evenly spread) + all %d read reductions + post-exposure chain, '
                       'sub-sample time scaled by %d/%d; %.1f s of CPU work' % (
                           len(picks), N, len(ri), N, len(picks), wall)),
            'seconds_per_exposure': per_exposure, 'photons_per_exposure': photons,
            'psf_alone': psf_info}
    if psf_info and 'apply_psf_cython_s_1thread' in psf_info:
        # what the real reference pays on top: its Cython wrapper's per-element copy loops
        # (pyparallel.pyx:23-25, 31-34), once per sub-sample
        extra = psf_info['apply_psf_cython_s_1thread'] - psf_info['psf_s_by_threads']['1']
        info['seconds_per_exposure_with_cython_wrapper'] = per_exposure + max(0.0, extra) * N
    return per_exposure, info


def cpu_baseline_object(val, info):
    out = {'value': val, 'unit': 'exposures/s', 'cores': info['cores'], 'kind': info['kind'],
           'sample': info['sample'], 'photons_per_s': info['photons_per_exposure'] * val,
           'note': 'lower bound on the reference time: numpy restatement without astropy unit algebra, '
                   'calibration FITS read once instead of per read, PSF() through ctypes'}
    if info.get('psf_alone'):
        pa = info['psf_alone']
        e = pa['electrons_per_call']
        out['native_kernel_alone'] = {
            'electrons_per_call': e, 'bins': pa['bins'], 'frame': pa['frame'],
            'Melectrons_per_s_by_threads': {k: round(e / v / 1e6, 2) for k, v in pa['psf_s_by_threads'].items()},
            'apply_psf_cython_Melectrons_per_s_1thread': (round(e / pa['apply_psf_cython_s_1thread'] / 1e6, 2)
                                                          if 'apply_psf_cython_s_1thread' in pa else None)}
    if 'seconds_per_exposure_with_cython_wrapper' in info:
        out['with_cython_wrapper'] = {
            'value': 1.0 / info['seconds_per_exposure_with_cython_wrapper'], 'unit': 'exposures/s',
            'note': 'the same exposure with apply_psf() (the unmodified Cython wrapper the reference '
                    'calls per sub-sample) instead of PSF() through ctypes'}
    return out


def run_reference(args, wk):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    calibration_dir(wk)
    inp = make_inputs(wk)
    from wayne import units as u
    N, W = len(np.asarray(u.value_in(inp['mid'], u.ms))), len(inp['wl'])
    R = len(inp['read_index'])
    # The first timed step is one FULL exposure (every sub-sample, ~30 s on this shape); the other
    # K - 1 steps are bounded samples of it (2-8 sub-samples per read interval, extrapolated), so that
    # the whole --steps K --warmup W run stays within a few minutes.  The line says what was sampled.
    n_sub = max(2, min(8, 64 // max(1, args.steps))) * R
    from oracle import psf as OP
    psf_info = time_reference_psf(psf_case_of(wk, inp), os.cpu_count() or 1) if OP.have_reference() else None
    times, info, full_info = [], None, None
    for i in range(args.warmup + args.steps):
        timed = i >= args.warmup
        want_full = timed and i == args.warmup and not args.no_full
        t, info = cpu_reference_exposure(wk, inp, None if want_full else (n_sub if timed else R),
                                         threads=info['cores'] if info else None, psf_alone=False)
        if want_full:
            full_info = dict(info, seconds=t)
        if timed:
            times.append(t)
    sec = float(np.mean(times))
    val = 1.0 / sec
    info['psf_alone'] = psf_info
    if psf_info and 'apply_psf_cython_s_1thread' in psf_info:
        info['seconds_per_exposure_with_cython_wrapper'] = sec + max(
            0.0, psf_info['apply_psf_cython_s_1thread'] - psf_info['psf_s_by_threads']['1']) * N
    if full_info:
        info['sample'] = '%d timed step(s): step 1 = %s' % (len(times), full_info['sample'])
        if len(times) > 1:
            info['sample'] += ('; steps 2..%d = bounded samples (%d sub-samples each + all read reductions + '
                               'post-exposure chain, sub-sample time scaled to %d)' % (len(times), n_sub, N))
        info['photons_per_exposure'] = full_info['photons_per_exposure']
    line = {'impl': 'reference', 'metric': 'exposures_per_s', 'value': val, 'unit': 'exposures/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'photons_per_s': info['photons_per_exposure'] / sec,
            'config': bench_config(wk, N, W), 'rng': 'numpy + rand_r (the reference\'s streams)',
            'step_seconds': [round(t, 3) for t in times],
            'full_exposure_seconds': round(full_info['seconds'], 3) if full_info else None,
            'cpu_baseline': cpu_baseline_object(val, info),
            'e2e': {'value': val, 'unit': 'exposures/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------
def copy_ceiling_for(world):
    """The box's measured copy ceiling for this GPU count (tools/copy_ceiling.py, committed under
    profiles/), or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_copy_ceiling.json')) as f:
            return json.load(f).get(str(world))
    except Exception:      # noqa: BLE001
        return None


def run_multi_visit(wk, inp, wl_q, local, rank, world, barrier, max_over_ranks, n_visits=8, n_exp=128):
    """BASELINE configs[4]: n_visits x n_exp exposures of the workload's shape, pointing random walk
    sigma = 0.02 px per exposure, partitioned exposure-wise over the ranks (sharding.run_sharded; no
    data-path collective), reads copied to the host, per-exposure summaries gathered.  The reference's
    counterpart is the serial loop of Observation.run_observation (wayne/observation.py:403-405)."""
    import torch
    import torch.distributed as dist
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    from wayne_b200 import sharding
    from wayne_b200.lightcurve import SeparableSignal
    total = n_visits * n_exp
    # per visit: a random walk of the pointing and a transit-like light curve sliding through the visit
    walk = np.empty((n_visits, n_exp, 2))
    for v in range(n_visits):
        g = np.random.default_rng(4242 + v)
        walk[v] = np.cumsum(0.02 * g.standard_normal((n_exp, 2)), axis=0)
    nsub = len(inp['lightcurve'])
    group = dist.new_group(backend='gloo') if world > 1 else None     # host-side metadata only

    def make(i, key):
        v, k = divmod(i, n_exp)
        lc = 0.5 * (1 + np.tanh((np.linspace(-1, 1, nsub) + 2.0 * (k / float(n_exp) - 0.5)) * 3))
        eg = ExposureGenerator(*inp['eg_args'], filename='v%02d_%04d_raw.fits' % (v, k + 1), rng='philox',
                               device=local)
        kw = frame_kwargs(wk, i)
        kw.update(x_ref=wk['x_ref'] + walk[v, k, 0], y_ref=wk['y_ref'] + walk[v, k, 1])
        exp = eg.scanning_frame(kw.pop('x_ref'), kw.pop('y_ref'), kw.pop('x_jitter'), kw.pop('y_jitter'),
                                wl_q, inp['flux'], SeparableSignal(lc, inp['depth0']),
                                kw.pop('scan_speed'), kw.pop('sample_rate'), inp['mid'], inp['dur'],
                                inp['read_index'], rng_key=(1963 + v, key[1]), **kw)
        return eg, exp

    def finish(h):
        eg, exp = h
        last = exp.reads[-1][0]                      # waits for the device->host copy
        return {'rank': rank, 'photons': eg.photons, 'last_read_sum': float(last[::64, ::64].sum())}

    # warm-up on a few exposures of the batch, then the whole batch
    sharding.run_sharded(min(total, 4 * world), make, world, rank, visit_seed=0, group=group, finish=finish,
                         pipeline_depth=3)
    barrier()
    import gc
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    merged = sharding.run_sharded(total, make, world, rank, visit_seed=0, group=group, finish=finish,
                                  pipeline_depth=3)
    torch.cuda.synchronize()
    sec = max_over_ranks(time.perf_counter() - t0)
    gc.enable()
    assert sorted(merged) == list(range(total))
    per_rank = [sum(1 for m in merged.values() if m['rank'] == r) for r in range(world)]
    return {'config': 'BASELINE configs[4]: %d visits x %d exposures of this workload\'s shape, pointing random '
                      'walk sigma = 0.02 px, exposure-wise over %d rank(s) through sharding.run_sharded, planet '
                      'signal per exposure (SeparableSignal), reads float64 to the host, summaries gathered '
                      '(gloo)' % (n_visits, n_exp, world),
            'exposures': total, 'seconds': sec, 'value': total / sec, 'unit': 'exposures/s', 'scaling': 'strong',
            'exposures_per_rank': per_rank,
            'photons_total': int(sum(m['photons'] for m in merged.values())),
            'checksum': float(sum(m['last_read_sum'] for m in merged.values()))}


def run_psf_dropin(wk, inp):
    """The PSF() / apply_psf() drop-in (include/wayne_b200.h; wayne/pyparallel.pyx:14-38 ->
    wayne/pyparallel_menu.c:10-113) timed at one sub-sample's shape with HOST arrays in and out, next
    to the reference's own C kernel and Cython wrapper on this host; outputs compared bit for bit."""
    from oracle import psf as OP
    from wayne_b200 import pyparallel
    ncpu = os.cpu_count() or 1
    out = []
    c1 = WORKLOADS['c1']
    inp1 = inp if wk is c1 else make_inputs(c1)
    for tag, w_, i_, electrons in (('configs[0] sub-sample: 4494 bins, 5e5 electrons, 256 x 256', c1, inp1, 5.0e5),
                                   ('1e7 electrons on 1014 x 1014', WORKLOADS['c4'], inp if wk is WORKLOADS['c4']
                                    else make_inputs(WORKLOADS['c4']), 1.0e7)):
        case = psf_case_of(w_, i_, electrons=electrons)
        cnt, xs, ys, ratio, sigl, sigh, L_ = case

        def best_of(fn, n=7):
            fn()
            b = None
            for _ in range(n):
                t0 = time.perf_counter()
                r = fn()
                el = time.perf_counter() - t0
                b = el if b is None or el < b else b
            return b, r

        t_psf, frame = best_of(lambda: pyparallel.psf_frame(cnt, xs, ys, ratio, sigl, sigh, L_, L_, 7, 2))
        t_apply, flat = best_of(lambda: pyparallel.apply_psf(cnt, xs, ys, ratio, sigl, sigh, L_, L_, 7, 2))
        row = {'case': tag, 'electrons': int(cnt.sum()), 'bins': int(len(cnt)), 'frame': int(L_),
               'PSF_ms': t_psf * 1e3, 'apply_psf_ms': t_apply * 1e3,
               'Melectrons_per_s': cnt.sum() / t_psf / 1e6}
        if OP.have_reference():
            ref = time_reference_psf(case, ncpu)
            best_t = min(ref['psf_s_by_threads'].values())
            row['reference'] = {'PSF_ms_by_threads': {k: v * 1e3 for k, v in ref['psf_s_by_threads'].items()},
                                'apply_psf_cython_ms_1thread': ref.get('apply_psf_cython_s_1thread', 0) * 1e3 or None}
            row['speedup_vs_reference_PSF_best_threads'] = best_t / t_psf
            if ref.get('apply_psf_cython_s_1thread'):
                row['speedup_vs_reference_apply_psf'] = ref['apply_psf_cython_s_1thread'] / t_apply
            want = OP.psf_reference(cnt, xs, ys, ratio, sigl, sigh, L_, L_, 7, 2)
            row['bit_exact_vs_reference'] = bool(np.array_equal(frame, want) and
                                                 np.array_equal(flat.reshape(L_, L_), want.astype(np.float64)))
            if not row['bit_exact_vs_reference']:
                raise RuntimeError("PSF drop-in differs from the reference kernel on the bench case")
        out.append(row)
    return out


def run_native(args, wk):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    from wayne_b200.engine import bind_to_gpu_numa_node
    numa_bound = bind_to_gpu_numa_node(local) if world > 1 else False
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # stdout carries the JSON line only
        dist.init_process_group('nccl', device_id=dev)
    calibration_dir(wk)
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    from wayne_b200 import _lib
    from wayne_b200.engine import DeviceEngine
    eng = DeviceEngine.get(local)
    if os.environ.get('WB200_HOSTTRACE'):
        from tools import hosttrace
        hosttrace.install()
    inp = make_inputs(wk)
    N, W = len(inp['read_index']) and len(np.asarray(u.value_in(inp['mid'], u.ms))), len(inp['wl'])

    # host inputs in pinned memory (e2e leg) and a device-resident copy (value leg)
    depth_pin = torch.empty((N, W), dtype=torch.float64, pin_memory=True)
    depth_host = depth_pin.numpy()
    depth_host[:] = inp['depth0'][None, :] * inp['lightcurve'][:, None]
    depth_dev = depth_pin.to(dev)
    wl_q = inp['wl'] * u.micron

    # the same planet signal in the visit driver's form (wayne_b200.lightcurve.ChebyshevSignal):
    # depth[s][w] = lightcurve[s] * depth0[w] is linear in x[w] = normalised depth0[w]
    from wayne_b200.lightcurve import ChebyshevSignal
    d0 = inp['depth0']
    mid0, half0 = 0.5 * (d0.max() + d0.min()), 0.5 * (d0.max() - d0.min())
    cheb_signal = ChebyshevSignal(np.c_[inp['lightcurve'] * mid0, inp['lightcurve'] * half0], (d0 - mid0) / half0)
    # ... and as its two factors (bit-identical to the dense product)
    from wayne_b200.lightcurve import SeparableSignal
    sep_signal = SeparableSignal(inp['lightcurve'], inp['depth0'])

    def one(i, resident):
        eg = ExposureGenerator(*inp['eg_args'], filename='%04d_raw.fits' % (i + 1), rng='philox', device=local)
        kw = frame_kwargs(wk, rank * 100000 + i)
        signal = {True: depth_dev, 'driver': cheb_signal, 'separable': sep_signal}.get(resident, depth_host)
        resident = resident is True
        exp = eg.scanning_frame(kw.pop('x_ref'), kw.pop('y_ref'), kw.pop('x_jitter'), kw.pop('y_jitter'),
                                wl_q, inp['flux'], signal,
                                kw.pop('scan_speed'), kw.pop('sample_rate'), inp['mid'], inp['dur'],
                                inp['read_index'], rng_key=(1963, rank * 100000 + i),
                                device_result=resident, **kw)
        return eg, exp

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: device-resident inputs, CUDA events on the launching stream ----
    # (the NVML sampler thread is started before the warm-up so its start-up cost
    # is outside the timed region; its samples are reset when the clock starts)
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get('WB200_NO_CLOCKS')) else None
    eng.profile = False              # the warm-up runs like the timed region (see below)
    phot_acc = torch.zeros((), dtype=torch.int64, device=dev)
    lost_acc = torch.zeros((1,), dtype=torch.int64, device=dev)
    tally_acc = torch.zeros((2,), dtype=torch.int64, device=dev)
    n_warm = max(args.warmup, 8)     # long enough for the allocator pools to become stationary
    for i in range(n_warm):
        eg, _ = one(i, True)
        phot_acc += eg._run.thrown()                 # same bookkeeping as the timed loop
        if eg._run.lost is not None:
            lost_acc += eg._run.lost
        if eg._run.tally is not None:
            tally_acc += eg._run.tally
        del eg
    barrier()
    from wayne_b200.engine import tune_host
    tune_host()                      # gc.freeze(): no 40 ms full-GC pauses inside the timed regions
    if sampler:
        sampler.reset()
    # per-stage CUDA events are OFF inside the timed region: with them on the library keeps every
    # kernel of an exposure on one stream (wb200_exposure_run), without them consecutive exposures
    # overlap (the next exposure's tables and counts run beside this one's electron throw).  The
    # per-kernel durations are measured right after it, on the same inputs, one kernel at a time.
    eng.profile = False
    eng.stage_times()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream(dev)
    phot_acc.zero_()
    lost_acc.zero_()
    tally_acc.zero_()
    torch.cuda.synchronize(dev)
    if os.environ.get('WB200_HOSTTRACE'):
        sys.stderr.write('[hosttrace] t=%.1f VALUE REGION START\n' % (time.perf_counter() * 1e3 % 1e6))
    e0.record(st)
    geom = None
    v_issue = []
    for i in range(args.steps):
        ta = time.perf_counter()
        eg, _ = one(n_warm + i, True)
        v_issue.append((time.perf_counter() - ta) * 1e3)
        phot_acc += eg._run.thrown()             # device-side bookkeeping, no sync
        if eg._run.lost is not None:
            lost_acc += eg._run.lost
        if eg._run.tally is not None:
            tally_acc += eg._run.tally           # [binned inside the frame, dropped outside it]
        geom = eg._run.win_geometry
        del eg
        if os.environ.get('WB200_HOSTTRACE') and (time.perf_counter() - ta) * 1e3 > 8:
            sys.stderr.write('[hosttrace] step %d: one() %.2f ms, whole iteration %.2f ms\n' % (
                i, v_issue[-1], (time.perf_counter() - ta) * 1e3))
    e1.record(st)
    barrier()
    launches = _lib.launch_count() - l0
    if os.environ.get('WB200_HOSTTRACE'):
        sys.stderr.write('[hosttrace] t=%.1f VALUE REGION END\n' % (time.perf_counter() * 1e3 % 1e6))
    ms_value = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    eng.profile = True
    eng.stage_times()
    n_prof = max(4, min(args.steps, 12))
    for i in range(n_prof):
        eg, _ = one(n_warm + args.steps + i, True)
        del eg
    stages = eng.stage_times()
    eng.profile = False
    clocks = sampler.stop() if sampler else None
    photons = float(phot_acc.item()) / args.steps
    if int(lost_acc.item()):
        raise RuntimeError("electrons fell outside their sub-sample windows")
    # electron conservation over the timed region, exact: every electron drawn was either
    # binned into a read-interval plane or dropped outside the frame (pyparallel_menu.c:93)
    binned, dropped = (int(v) for v in tally_acc.cpu().numpy())
    thrown = int(phot_acc.item())
    from wayne_b200 import params as _params
    if _params.direct_accumulation and (binned + dropped != thrown or binned <= 0):
        raise RuntimeError("electron bookkeeping of the timed exposures is off: %d binned + %d dropped != "
                           "%d thrown" % (binned, dropped, thrown))

    # ---- e2e: host buffers through the public API, H2D + D2H inside ---------------
    import collections

    def pipeline(first, n, mode=False):
        """n exposures through the public API; every exposure's reads are touched
        on the host (three exposures behind the one being issued)."""
        pending = collections.deque()
        t_issue, t_wait, d2h, checksum = [], [], 0, 0.0
        for i in range(n):
            ta = time.perf_counter()
            _, exp = one(first + i, mode)
            tb = time.perf_counter()
            pending.append(exp)
            while len(pending) > 3 or (i == n - 1 and pending):
                reads = pending.popleft().reads
                d2h = sum(r[0].nbytes for r in reads)
                checksum += float(reads[-1][0][512 % reads[-1][0].shape[0], 7])
            t_issue.append((tb - ta) * 1e3)
            t_wait.append((time.perf_counter() - tb) * 1e3)
        torch.cuda.synchronize(dev)
        return t_issue, t_wait, d2h, checksum

    pipeline(0, max(12, args.warmup))        # warm-up: same pipeline, fills the pinned / device pools
    barrier()
    if os.environ.get('WB200_HOSTTRACE'):
        sys.stderr.write('[hosttrace] t=%.1f E2E REGION START\n' % (time.perf_counter() * 1e3 % 1e6))
    import gc
    gc.collect()
    gc.disable()                     # reference counting frees the exposures; no cyclic-GC pause in the timed legs
    t0 = time.perf_counter()
    t_issue, t_wait, d2h, checksum = pipeline(n_warm, args.steps)
    torch.cuda.synchronize(dev)
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    gc.enable()
    h2d = depth_host.nbytes + inp['flux'].nbytes + inp['wl'].nbytes + 3 * 8 * N
    # the visit driver's form of the same exposure: planet signal as a Chebyshev expansion
    pipeline(0, max(12, args.warmup), 'driver')   # same warm-up as the e2e leg (a short one left a cudaMalloc in the timed region)
    barrier()
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    d_issue, d_wait, _, _ = pipeline(n_warm, args.steps, 'driver')
    ms_drv = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    gc.enable()
    h2d_drv = cheb_signal.coef.nbytes + cheb_signal.x.nbytes + inp['flux'].nbytes + inp['wl'].nbytes + 3 * 8 * N
    # the planet signal as its two factors (lightcurve.SeparableSignal)
    pipeline(0, max(12, args.warmup), 'separable')
    barrier()
    gc.collect()
    gc.disable()
    t0 = time.perf_counter()
    s_issue, s_wait, _, _ = pipeline(n_warm, args.steps, 'separable')
    ms_sep = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    gc.enable()
    h2d_sep = 8 * (N + W) + inp['flux'].nbytes + inp['wl'].nbytes + 3 * 8 * N

    # ---- what THIS box can move: concurrent pinned copies of the same sizes on the same GPUs ----
    # (tools/copy_ceiling.py; a property of the host -- PCIe, the VM's memory path -- that the e2e
    # legs above cannot exceed however fast the kernels are; measured here, on this box, in this run,
    # because it differs from box to box: profiles/r02_copy_ceiling.json holds one box's 1/2/4/8 sweep)
    ceiling = None
    if not os.environ.get('WB200_NO_CEILING'):
        try:
            from tools.copy_ceiling import measure as copy_ceiling_measure
            ceiling = copy_ceiling_measure(dev, world, 0.6, d2h_bytes=d2h, h2d_bytes=depth_host.nbytes)
            ceiling.update(n_gpus=world, seconds_per_mode=0.6, d2h_bytes=int(d2h), h2d_bytes=int(depth_host.nbytes),
                           source='tools/copy_ceiling.measure, this run, this box, all ranks at once')
        except Exception as exc:      # noqa: BLE001
            ceiling = None
            sys.stderr.write('copy ceiling not measured: %r\n' % (exc,))

    # ---- BASELINE configs[4]: the multi-visit batch, exposure-wise over the ranks -------------
    multi_visit = None
    if not args.no_extras:
        multi_visit = run_multi_visit(wk, inp, wl_q, local, rank, world, barrier, max_over_ranks,
                                      n_visits=(8 if wk['photons'] >= 1e8 else 2),
                                      n_exp=(128 if wk['photons'] >= 1e8 else 4))
    # per-kernel durations of the driver form (its k_counts evaluates the Chebyshev signal)
    eng.profile = True
    eng.stage_times()
    for i in range(4):
        one(n_warm + i, 'driver')
    stages_drv = eng.stage_times()
    eng.profile = False

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the per-kernel durations measured above ------------------------
    S = wk['sub']
    F = min(S + 10, 1024)
    R = wk['nsamp'] - 1
    per = lambda name: (stages[name][0] / max(1, stages[name][1])) if name in stages else None   # noqa: E731
    hbm_peak, peak_src = measured_peaks()
    from wayne_b200 import params as _p
    pb = 4 if (_p.use_context and _p.direct_accumulation) else 8    # resident planes: float32 through the context
    reads_bytes = (R * F * F * 8            # interval accumulators (int64 fixed point) read ...
                   + (R * F * F * 8 if pb == 4 else 0)   # ... and written back as zeros (no memset pass)
                   + 2 * R * F * F * pb     # dark, dark error
                   + 2 * F * F * pb         # sky, gain
                   + 4 * F * F * pb         # non-linearity planes (1+c1, c2, c3, c4; the derivative's
                                            # coefficients are formed from them in the native kernel)
                   + (F * F * 8 if S == 256 else 0) + F * F * 4   # zero read, cosmic heads
                   + (R + 1) * F * F * 8)   # NSAMP reads written (float64 like the reference)
    t_reads, t_throw, t_gather, t_counts = per('k_reads'), per('k_throw'), per('k_gather'), per('k_counts')
    ww, wh, chunk = geom
    gather_bytes = (N * ww * wh * 4 + 2 * R * F * F * 8) if ww else 0     # only on the window/gather (parity) path
    roof_hbm = {'kernel': 'k_reads_native', 'bound': 'hbm', 'achieved': reads_bytes / (t_reads * 1e-3) / 1e9,
                'peak': hbm_peak, 'unit': 'GB/s', 'peak_source': peak_src,
                'frac': reads_bytes / (t_reads * 1e-3) / 1e9 / hbm_peak, 'traffic': None,
                'algorithmic_bytes': reads_bytes, 'ms': t_reads}
    # measured denominators for the photon kernel (no memory traffic to speak of):
    #   mb(7)  the thrower's random recipe and nothing else (one Philox4x32-10 call + four
    #          fp32 SFU Box-Muller pairs) -- the ceiling of GENERATING electrons on this GPU
    #   mb(6)  the Philox4x32-10 calls alone; mb(8) IMAD.WIDE.U32 alone (quarter rate on B200:
    #          20 of them bound a call at ~80 cycles per warp)
    #   mb(2)  shared-memory atomics with the PSF's 3x3 same-address pattern
    #   mb(1)  conflict-free shared-memory atomics
    import ctypes as C

    def microbench(which, iters):
        ms, ops = C.c_double(), C.c_double()
        _lib.check(_lib.lib.wb200_microbench(which, iters, C.byref(ms), C.byref(ops)), 'microbench')
        return ops.value / (ms.value * 1e-3) / 1e9

    try:
        rng_peak, atom_psf, atom_free = microbench(7, 2048), microbench(2, 4096), microbench(1, 4096)
        philox_calls, imad_wide = microbench(6, 2048), microbench(8, 2048)
    except Exception:      # noqa: BLE001
        rng_peak = atom_psf = atom_free = philox_calls = imad_wide = None
    traffic = {}
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')) as f:
            traffic = json.load(f)
    except Exception:      # noqa: BLE001
        pass
    roof_hbm['traffic'] = traffic.get('k_reads')
    roof_hbm['traffic_source'] = traffic.get('source')
    roof_throw = {'kernel': 'k_throw_philox (direct accumulation)',
                  'bound': 'sm_issue: one Philox4x32-10 call per four electrons (IMAD.WIDE pipe) + SFU Box-Muller + '
                           'shared-memory atomic per electron (HBM idle)',
                  'achieved': photons / (t_throw * 1e-3) / 1e9, 'peak': rng_peak,
                  'unit': 'Gelectron/s',
                  'peak_source': 'wb200_microbench(7): the thrower\'s random recipe only (Philox4x32-10 call + 4 fp32 '
                                 'SFU Box-Muller pairs), all SMs, this run',
                  'philox_calls_gcalls_s': philox_calls, 'imad_wide_gops_s': imad_wide,
                  'frac': (photons / (t_throw * 1e-3) / 1e9 / rng_peak) if rng_peak else None,
                  'smem_atomic_peaks_gops': {'psf_like_3x3': atom_psf, 'conflict_free': atom_free},
                  # one shared-memory increment per in-frame electron
                  'increments_over_atomic_peak': ((photons / (t_throw * 1e-3) / 1e9 / atom_psf) if atom_psf else None),
                  'mufu_bound_gelectron_s': (148 * 16 * (clocks['sm_mhz'] if clocks and clocks.get('sm_mhz')
                                                        else 1965.0) * 1e6 / 4) / 1e9,
                  'traffic': traffic.get('k_throw'), 'traffic_source': traffic.get('source'), 'ms': t_throw}
    # the same achieved rate against the hardware bound that does not depend on the recipe's own
    # microbenchmark: four MUFU per electron at one warp-instruction per 8 cycles per sub-partition
    roof_throw['frac_of_mufu_bound'] = roof_throw['achieved'] / roof_throw['mufu_bound_gelectron_s']
    dominant = max(((k, v) for k, v in stages.items() if k.startswith('k_')), key=lambda kv: kv[1][0])[0]
    if ceiling is None:
        ceiling = copy_ceiling_for(world)      # a committed sweep of another box, as a fallback
    e2e_val = world * 1e3 / ms_e2e

    def frac_of(key, v):
        c = (ceiling or {}).get('ceiling_exposures_per_s', {})
        for k, cv in c.items():
            if k.startswith(key) and cv:
                return v / cv
        return None

    line = {
        'metric': 'exposures_per_s', 'value': world * 1e3 / ms_value, 'unit': 'exposures/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'warmup_performed': n_warm,
        'ms_per_step': ms_value,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'photons_per_exposure': photons,
        'electron_bookkeeping': {'thrown': thrown, 'binned_in_frame': binned, 'dropped_off_frame': dropped,
                                 'check': 'binned + dropped == thrown over the timed exposures (asserted)'},
        'photons_per_s': world * photons * 1e3 / ms_value,
        'config': bench_config(wk, N, W), 'rng': 'philox (native streams)', 'chunk_bins': chunk,
        'clocks': clocks,
        'e2e': {'value': e2e_val, 'unit': 'exposures/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'form': 'planet signal = dense HOST array [n_samples][n_bins] (the reference\'s argument, '
                        'uploaded every exposure); reads float64 to the host',
                'host_issue_ms': [round(float(np.median(t_issue)), 3), round(float(np.max(t_issue)), 3)],
                'host_wait_ms': [round(float(np.median(t_wait)), 3), round(float(np.max(t_wait)), 3)],
                'frac_of_copy_ceiling': frac_of('dense_host_signal', e2e_val),
                'driver_form': {
                    'value': world * 1e3 / ms_drv, 'unit': 'exposures/s', 'ms_per_step': ms_drv,
                    'h2d_bytes_per_step': int(h2d_drv), 'd2h_bytes_per_step': int(d2h),
                    'host_issue_ms': [round(float(np.median(d_issue)), 3), round(float(np.max(d_issue)), 3)],
                    'frac_of_copy_ceiling': frac_of('factored_signal', world * 1e3 / ms_drv),
                    'note': 'the same exposure with the planet signal in the visit driver\'s form '
                            '(lightcurve.ChebyshevSignal, evaluated inside k_counts): what every exposure of '
                            'Observation.run_observation does; no 135 MB upload'},
                'separable_form': {
                    'value': world * 1e3 / ms_sep, 'unit': 'exposures/s', 'ms_per_step': ms_sep,
                    'h2d_bytes_per_step': int(h2d_sep), 'd2h_bytes_per_step': int(d2h),
                    'host_issue_ms': [round(float(np.median(s_issue)), 3), round(float(np.max(s_issue)), 3)],
                    'frac_of_copy_ceiling': frac_of('factored_signal', world * 1e3 / ms_sep),
                    'note': 'planet signal handed over as its two factors (lightcurve.SeparableSignal: '
                            'lightcurve[n_samples], depth[n_bins]); bit-identical frames to the dense array'},
                'copy_ceiling': ceiling},
        'multi_visit': multi_visit,
        'gpu_launches': int(launches), 'numa_bound': bool(numa_bound),
        'host_issue_ms': [round(float(np.median(v_issue)), 3), round(float(np.max(v_issue)), 3)],
        'stage_ms': {k: v[0] / max(1, v[1]) for k, v in stages.items()},
        'stage_ms_note': 'one kernel at a time (per-stage events keep an exposure on one stream), %d exposures '
                         'right after the timed region; inside it consecutive exposures overlap' % n_prof,
        'stage_ms_driver': {k: v[0] / max(1, v[1]) for k, v in stages_drv.items()},
        'dominant_kernel': dominant,
        'roofline': roof_throw if dominant == 'k_throw' else roof_hbm,
        'roofline_hbm': roof_hbm,
        'roofline_photons': roof_throw,
        'gather': ({'ms': t_gather, 'algorithmic_bytes': gather_bytes,
                    'achieved_gbs': gather_bytes / (t_gather * 1e-3) / 1e9} if t_gather else None),
    }
    if world == 1 and not args.no_extras:
        line['psf_dropin'] = run_psf_dropin(wk, inp)
    if world == 1 and not args.no_cpu:
        sec, info = cpu_reference_exposure(wk, inp, 48 * len(inp['read_index']))   # ~10-15 s of CPU work
        line['cpu_baseline'] = cpu_baseline_object(1.0 / sec, info)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=('native', 'reference'))
    ap.add_argument('--workload', default='c4', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-extras', action='store_true', help='skip the multi-visit and PSF drop-in legs')
    ap.add_argument('--no-full', action='store_true',
                    help='reference arm: bounded samples only (no full un-sampled exposure as step 1)')
    args = ap.parse_args()
    wk = WORKLOADS[args.workload]
    if args.impl == 'reference':
        run_reference(args, wk)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_native(args, wk)


if __name__ == '__main__':
    main()
