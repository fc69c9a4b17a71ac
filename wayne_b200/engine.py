"""Device engine: one exposure = one batched job on one B200.

This is the host side of the hot path.  It owns device buffers as torch tensors
(torch is used for nothing else), keeps the calibration planes resident in HBM
per (grism, SUBARRAY, mode), and drives the stage kernels of libwayne_b200.so
through ctypes, all on one CUDA stream with no host synchronisation until the
result is copied back.

Replaces the reference's serial per-sub-sample loop
(wayne/exposure_generator.py:336-394) and its per-read / post-exposure numpy
passes (:361-389, :407-444).  Policy (argument handling, RNG order of the
compat mode, units) lives in exposure_generator.py; this module is mechanism.

HBM layout (all row-major, float64 unless noted)
  tables     wl, ratio, sigl, sigh, sens, dwl           [W]
  trace      {x_ref,y_ref,m_t,c_t,m_w,c_w,m_wl,c_wl}    [N][8]
  depth      planet signal (array form only)            [N][ld], ring of upload buffers
  counts     int32                                      [N][W]
  acc        electrons per read interval, bordered      [R][F][F]  int64 2^-24 fixed point
                                                        (native) / float64 (parity path)
  windows    int32, parity path only                    [nb][WH][WW] one per sub-sample
  planes     sky, gain, zero, nl[7], dark/err[R]        [F][F] each (bordered), resident
  out        the NSAMP reads                            [R+1][F][F] -> pooled pinned host memory

Streams: compute (torch's current stream), upload, download; see DESIGN.md 6b.
"""
from __future__ import annotations

import ctypes as C
import os
import math
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import lib, check

BORDER = 5
TILE_W, TILE_H = 128, 64            # must match csrc/wayne_b200.cu
# largest |z| a stream can produce: the native thrower reaches sqrt(2*49*ln 2) = 8.24 (photons.cuh), rand_r 6.6
ZMAX = {_lib.RNG_PHILOX: 8.3, _lib.RNG_RANDR: 6.6}
WINDOW_BYTES_CAP = 2 << 30          # HBM spent on sub-sample windows per batch


def _lib_params_calb():
    from . import params
    return params._calb_dir


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _dp(a):
    return a.ctypes.data_as(_lib.DP)


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs closest to its GPU (NVML's ideal CPU
    affinity), so pinned host buffers are first-touched on the GPU's NUMA node
    and per-rank copies do not cross sockets.  Best effort: returns False when
    NVML is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = device_index
        if vis and vis.split(',')[device_index].strip().isdigit():
            phys = int(vis.split(',')[device_index])
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        return True
    except Exception:      # noqa: BLE001
        return False


def tune_host():
    """Call once after start-up in a long-running driver: moves every object alive
    now into the garbage collector's permanent generation (gc.freeze), so the
    cyclic collector's full passes -- 40 ms each with torch imported, enough to
    let the GPU run dry -- no longer rescan them while exposures are in flight."""
    import gc
    gc.collect()
    gc.freeze()


class DeviceEngine(object):
    """Per-GPU context: stream, resident calibration planes, scratch reuse."""

    _engines = {}

    @classmethod
    def get(cls, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError(
                "wayne_b200 needs a CUDA device: the exposure path has no CPU implementation "
                "(the CPU restatement under oracle/ is test infrastructure)")
        if device is None:
            device = torch.cuda.current_device()
        device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in cls._engines:
            cls._engines[key] = cls(torch.device("cuda", key))
        return cls._engines[key]

    def __init__(self, device):
        self.device = device
        self._planes = {}
        self._pinned = {}
        self._pinned_made = {}
        self._marks = []
        self._open = {}
        self._copy_stream = None
        self._upload_stream = None
        self._in_flight = []
        self._upload_rings = {}

    # -- helpers -----------------------------------------------------------
    def stream_ptr(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def to_dev(self, a, dtype=None):
        """Host array -> device tensor on the current stream (device tensors
        pass through, so callers may keep large inputs resident in HBM)."""
        if isinstance(a, torch.Tensor):
            if a.device != self.device:
                a = a.to(self.device, non_blocking=True)
            return a if a.is_contiguous() else a.contiguous()
        a = np.ascontiguousarray(a, dtype=dtype)
        return torch.from_numpy(a).to(self.device, non_blocking=True)

    _TORCH_DT = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                 np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}

    def to_dev_many(self, arrays, ahead=False):
        """Several small host arrays -> device tensors with ONE host->device copy:
        packed (16-byte aligned) into a pinned staging buffer and returned as
        typed views of the device buffer.  A dozen separate pageable copies per
        exposure cost ~35 us of stream time each.

        ahead=True issues the copy on the upload stream (the current stream waits
        for it): a copy in stream order on the compute stream is only handed to
        the copy engine when the previous exposure's kernels have finished, by
        which time the engine is busy with a later exposure's 135 MB upload and
        the small copy -- and with it the whole exposure -- waits behind it."""
        arrays = [np.ascontiguousarray(a) for a in arrays]
        offs, total = [], 0
        for a in arrays:
            if a.dtype not in self._TORCH_DT:
                raise TypeError("unsupported dtype {}".format(a.dtype))
            offs.append(total)
            total += (a.nbytes + 15) // 16 * 16
        stage = torch.empty((max(total, 16),), dtype=torch.uint8, pin_memory=True)
        host = stage.numpy()
        for a, o in zip(arrays, offs):
            host[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
        if ahead:
            _, up = self._streams()
            cur = torch.cuda.current_stream(self.device)
            with torch.cuda.stream(up):
                dev = stage.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(up)
            cur.wait_event(ev)
            dev.record_stream(cur)
        else:
            dev = stage.to(self.device, non_blocking=True)
        out = []
        for a, o in zip(arrays, offs):
            t = dev[o:o + a.nbytes].view(self._TORCH_DT[a.dtype]).reshape(a.shape)
            out.append(t)
        return out

    def pinned_out(self, shape, dtype):
        """A pinned host buffer for the exposure's device->host copy, as a numpy
        array.  Buffers are pooled per (shape, dtype) and return to the pool when
        the array (and every view of it) has been garbage collected, so the reads
        an Exposure holds are never overwritten by a later exposure."""
        key = (tuple(shape), dtype)
        free = self._pinned.setdefault(key, [])
        if not free:
            # cudaHostAlloc of a 126 MB buffer takes ~100 ms: a pool that grows one buffer at a
            # time stalls a pipelined visit whenever the garbage collector returns a buffer a
            # little late.  The first exposure of a shape gets one buffer; the second one (the
            # caller is evidently repeating the shape) makes the pool as deep as the pipeline
            # can get -- exposures in flight plus the ones the caller still holds.
            made = self._pinned_made.get(key, 0)
            grow = 1 if made == 0 else max(1, self.PINNED_RESERVE - made)
            for _ in range(grow):
                free.append(torch.empty(shape, dtype=dtype, pin_memory=True))
            self._pinned_made[key] = made + grow
        t = free.pop()
        arr = t.numpy()
        weakref.finalize(arr, free.append, t)
        return t, arr

    def empty(self, shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype=torch.float64):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def bordered(self, plane, F, fill=0.0):
        """L x L host plane -> F x F float64 host plane with the 5-pixel border."""
        plane = np.asarray(plane)
        out = np.full((F, F), fill, dtype=np.float64)
        n = plane.shape[0]
        out[BORDER:BORDER + n, BORDER:BORDER + n] = plane
        return out

    # exposures in flight before the host waits (~0.5 GB of HBM each); 3 and 5 measured no different on e2e
    MAX_IN_FLIGHT = int(os.environ.get('WAYNE_B200_MAX_IN_FLIGHT', '4'))
    PINNED_RESERVE = 10    # depth of the pinned download pool once a shape repeats (see pinned_out)

    def _streams(self):
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._upload_stream = torch.cuda.Stream(self.device)
        return self._copy_stream, self._upload_stream

    def upload_async(self, host_array):
        """Large host array -> device on the upload stream, into a ring of
        persistent device buffers (one per exposure that may be in flight); the
        current stream waits for the copy, so the upload of exposure i+1 overlaps
        the kernels of exposure i.  Returns (tensor, slot); call
        release_upload(slot) once the last kernel reading the tensor is queued.

        A ring instead of the caching allocator: a freshly freed block is only
        handed out again after the compute stream has drained, which serialised
        the upload of the next exposure behind the current one."""
        _, up = self._streams()
        cur = torch.cuda.current_stream(self.device)
        key = (tuple(host_array.shape), str(host_array.dtype))
        ring = self._upload_rings.setdefault(key, {'bufs': [], 'free': [], 'next': 0})
        n_slots = self.MAX_IN_FLIGHT + 1
        slot = ring['next'] % n_slots
        ring['next'] += 1
        if slot >= len(ring['bufs']):
            ring['bufs'].append(torch.empty(host_array.shape, dtype=self._TORCH_DT[host_array.dtype],
                                            device=self.device))
            ring['free'].append(None)
        buf = ring['bufs'][slot]
        with torch.cuda.stream(up):
            if ring['free'][slot] is not None:
                up.wait_event(ring['free'][slot])      # previous user's kernels are done
            buf.copy_(torch.from_numpy(host_array), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(up)
        cur.wait_event(ev)
        return buf, (key, slot)

    def release_upload(self, handle):
        key, slot = handle
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._upload_rings[key]['free'][slot] = ev

    def admit(self):
        """Called at the start of an exposure: wait until fewer than MAX_IN_FLIGHT
        earlier exposures are still executing.  Bounds the host's run-ahead (and
        with it HBM and pinned-memory use: ~1 GB per exposure in flight) without
        ever letting the GPU run dry -- issuing an exposure costs the host less
        than executing one costs the GPU."""
        while len(self._in_flight) >= self.MAX_IN_FLIGHT:
            self._in_flight.pop(0).synchronize()

    def retire(self, event=None):
        """Called when an exposure's last operation has been queued."""
        if event is None:
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream(self.device))
        self._in_flight.append(event)

    def fetch_async(self, dev_tensor, small=None):
        """Queue the device->host copy of ``dev_tensor`` (and of an optional small
        tensor, e.g. a status word) on the copy stream behind everything already on
        the current stream.  Returns (event, numpy array in pooled pinned memory,
        small host tensor or None); both are valid once the event has completed."""
        cs, _ = self._streams()
        t, arr = self.pinned_out(tuple(dev_tensor.shape), dev_tensor.dtype)
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        cs.wait_event(ready)
        small_host = None
        with torch.cuda.stream(cs):
            t.copy_(dev_tensor, non_blocking=True)
            if small is not None:
                small_host = torch.empty(small.shape, dtype=small.dtype, pin_memory=True)
                small_host.copy_(small, non_blocking=True)
                small.record_stream(cs)
        dev_tensor.record_stream(cs)
        done = torch.cuda.Event()
        done.record(cs)
        self.retire(done)
        return done, arr, small_host

    def fetch(self, dev_tensor):
        """Device tensor -> numpy array backed by pooled pinned memory; waits for
        the stream, so everything queued before it has finished."""
        done, arr, _ = self.fetch_async(dev_tensor)
        done.synchronize()
        return arr

    def exposure_context(self, grism, detector, subarray, sampseq):
        """The resident context of one instrument configuration on this GPU (created once)."""
        key = ('ctx', grism.name, getattr(grism, 'flat_file_name', None), getattr(grism, 'sky_file_name', None),
               detector.gain_file_name, detector.non_linear_file_name, int(subarray), sampseq,
               _lib_params_calb())
        if key not in self._planes:
            self._planes[key] = ExposureContext(self, grism, detector, subarray, sampseq)
        return self._planes[key]

    def cached_plane(self, key, make):
        if key not in self._planes:
            self._planes[key] = make()
        return self._planes[key]

    def drop_planes(self):
        self._planes.clear()

    # -- per-stage device timing (bench / profiles) ----------------------------
    profile = False

    def mark(self, name, begin):
        """With ``profile`` on, bracket a stage with CUDA events on the current
        stream (no synchronisation); read the totals with stage_times()."""
        if not self.profile:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        if begin:
            self._open[name] = ev
        else:
            self._marks.append((name, self._open.pop(name), ev))

    def stage_times(self, reset=True):
        """{stage: (total ms, launches)} of the marks recorded so far (synchronises)."""
        torch.cuda.synchronize(self.device)
        out = {}
        for name, a, b in self._marks:
            t, n = out.get(name, (0.0, 0))
            out[name] = (t + a.elapsed_time(b), n + 1)
        if reset:
            self._marks = []
        for ctx in self._planes.values():          # exposures queued through resident contexts
            if isinstance(ctx, ExposureContext):
                for name, (t, n) in ctx.stage_times().items():
                    t0, n0 = out.get(name, (0.0, 0))
                    out[name] = (t0 + t, n0 + n)
        return out


class ExposureRun(object):
    """One exposure's device state, built stage by stage."""

    def __init__(self, engine, grism, subarray, wl_um, flux, depth, depth_col0, xr, yr, dur_ms,
                 scale, read_end, aux=None):
        """
        aux                  {name: small host array} uploaded with the other small
                             inputs in the exposure's FIRST host->device copy
                             (device views in ``self.aux``).  Small copies issued
                             later would queue on the copy engine behind the next
                             exposure's planet-signal upload and stall this
                             exposure's per-pixel pass.
        wl_um, flux [W]      cropped wavelength grid [micron] and stellar flux on it
        depth [N][>=W] / None planet signal rows; columns depth_col0.. are used
        xr, yr, dur_ms [N]   sub-sample reference positions [px] and durations [ms]
        read_end [R]         index of the last sub-sample of each read
        """
        self.e = engine
        self.grism = grism
        self.S = int(subarray)
        self.L = 1014 if self.S == 1024 else self.S
        self.F = min(self.S + 10, 1024)
        self.W = int(len(wl_um))
        self.N = int(len(xr))
        self.R = int(len(read_end))
        self.sub_scale = float(507 - self.S // 2)     # exposure_generator.py:630 (py2 int division)
        self.scale = 1.0 if scale is None else float(scale)
        self.wl_host = np.ascontiguousarray(wl_um, dtype=np.float64)
        self.xr_host = np.ascontiguousarray(xr, dtype=np.float64)
        self.yr_host = np.ascontiguousarray(yr, dtype=np.float64)
        self.read_end_host = np.ascontiguousarray(read_end, dtype=np.int32)
        if self.W < 2:
            raise ValueError("need at least two wavelength bins inside the grism limits")
        if self.N < 1 or self.N > 65535:
            raise ValueError("1..65535 sub-samples per exposure supported, got {}".format(self.N))

        e = engine
        st = e.stream_ptr()
        self.cheb_order = 0
        cheb = depth if hasattr(depth, 'coef') and hasattr(depth, 'x') else None
        if cheb is not None:
            # planet signal as a per-sub-sample Chebyshev expansion (lightcurve.ChebyshevSignal):
            # evaluated inside k_counts, never materialised
            depth = None
            self.cheb_order = int(cheb.coef.shape[1])
            coef = cheb.coef[:self.N]
            self._cheb_dev = coef.contiguous() if isinstance(coef, torch.Tensor) else None
            self._cheb_host = (None if self._cheb_dev is not None
                               else np.ascontiguousarray(coef, dtype=np.float64),
                               np.ascontiguousarray(cheb.x[depth_col0:depth_col0 + self.W], dtype=np.float64))
        flux_is_dev = isinstance(flux, torch.Tensor)
        small = [self.wl_host, self.xr_host, self.yr_host, np.ascontiguousarray(dur_ms, dtype=np.float64),
                 self.read_end_host]
        aux = dict(aux or {})
        if self.cheb_order:
            aux['_cheb_x'] = self._cheb_host[1]
            if self._cheb_dev is None:
                aux['_cheb_coef'] = self._cheb_host[0]
        aux_names = sorted(aux)
        small += [aux[k] for k in aux_names]
        if not flux_is_dev:
            small.append(np.ascontiguousarray(flux, dtype=np.float64))
        packed = e.to_dev_many(small, ahead=True)
        self.d_wl, self.d_xr, self.d_yr, self.d_dur, self.d_read_end = packed[:5]
        self.aux = dict(zip(aux_names, packed[5:5 + len(aux_names)]))
        if self.cheb_order and self._cheb_dev is not None:
            self.aux['_cheb_coef'] = self._cheb_dev
        self.d_flux = e.to_dev(flux) if flux_is_dev else packed[-1]
        if depth is not None:
            if not isinstance(depth, torch.Tensor):
                depth = np.asarray(depth)
                if depth.dtype != np.float64 or not depth.flags.c_contiguous:
                    depth = np.ascontiguousarray(depth, dtype=np.float64)
            elif depth.dtype != torch.float64:
                raise ValueError("device planet_signal must be float64")
            big = (not isinstance(depth, torch.Tensor)) and depth.nbytes >= (8 << 20)
            self._depth_slot = None
            if big:
                self.d_depth_full, self._depth_slot = e.upload_async(depth)
            else:
                self.d_depth_full = e.to_dev(depth)
            self.depth_ld = int(depth.shape[1])
            self.depth_ptr = C.c_void_p(self.d_depth_full.data_ptr() + 8 * int(depth_col0))
        else:
            self.d_depth_full, self.depth_ld, self.depth_ptr = None, 0, None
            self._depth_slot = None

        # ---- stage 1a: wavelength-only tables ---------------------------------
        tabs = e.empty((5, self.W))
        self.d_ratio, self.d_sigl, self.d_sigh, self.d_sens, self.d_dwl = tabs
        poly = np.concatenate([np.asarray(grism.psf_ratio_poly.coeffs, dtype=np.float64),
                               np.asarray(grism.psf_sigmal_poly.coeffs, dtype=np.float64),
                               np.asarray(grism.psf_sigmah_poly.coeffs, dtype=np.float64)])
        assert poly.shape == (12,)
        self._poly = poly
        sens_wl, sens_val = grism._load_sens()
        key = ("sens", grism.name)
        d_swl, d_sval = e.cached_plane(key, lambda: (e.to_dev(sens_wl), e.to_dev(sens_val)))
        check(lib.wb200_bin_tables(self.W, _ptr(self.d_wl), _dp(poly), len(sens_wl), _ptr(d_swl),
                                   _ptr(d_sval), _ptr(self.d_ratio), _ptr(self.d_sigl),
                                   _ptr(self.d_sigh), _ptr(self.d_sens), _ptr(self.d_dwl), st),
              "wb200_bin_tables")
        # ---- stage 1b: trace per sub-sample -------------------------------------
        self.d_trace = e.empty((self.N, _lib.TRACE_STRIDE))
        a9 = np.asarray(grism.trace_coeff, dtype=np.float64)
        b9 = np.asarray(grism.wl_solution, dtype=np.float64)
        check(lib.wb200_trace_table(self.N, _ptr(self.d_xr), _ptr(self.d_yr), _dp(a9), _dp(b9),
                                    _ptr(self.d_trace), st), "wb200_trace_table")
        self.d_counts = None
        self.d_totals = e.empty((self.N,), torch.int64)
        self.d_expected = None
        self.d_acc = None
        self.lost = e.zeros((1,), torch.int64)
        self.tally = None                  # direct path: [binned in frame, dropped outside it]

    def _release_depth(self):
        """The planet-signal upload buffer may be overwritten by a later exposure
        once the kernels queued so far have run (k_counts is its only reader)."""
        if self._depth_slot is not None:
            self.e.release_upload(self._depth_slot)
            self._depth_slot = None
            self.d_depth_full = None
            self.depth_ptr = None
            self._depth_gone = True

    def __del__(self):
        try:
            self._release_depth()
        except Exception:      # interpreter shutdown
            pass

    # ------------------------------------------------------------------
    def tables_host(self):
        """(ratio, sigl, sigh, sens, dwl) as numpy -- parity tests."""
        return tuple(t.cpu().numpy() for t in
                     (self.d_ratio, self.d_sigl, self.d_sigh, self.d_sens, self.d_dwl))

    def trace_host(self):
        return self.d_trace.cpu().numpy()

    def positions_host(self):
        e = self.e
        xp = e.empty((self.N, self.W))
        yp = e.empty((self.N, self.W))
        check(lib.wb200_trace_positions(self.N, self.W, _ptr(self.d_trace), _ptr(self.d_wl),
                                        self.sub_scale, _ptr(xp), _ptr(yp), e.stream_ptr()),
              "wb200_trace_positions")
        return xp.cpu().numpy(), yp.cpu().numpy()

    def _counts_args(self, mode, key, d_expected, d_counts):
        a = _lib.CountsArgs()
        a.n_samples, a.n_bins, a.count_mode, a.cheb_order = self.N, self.W, mode, self.cheb_order
        a.key0, a.key1 = key[0] & 0xffffffff, key[1] & 0xffffffff
        a.scale, a.depth_ld = self.scale, self.depth_ld
        a.d_flux = self.d_flux.data_ptr()
        a.d_depth = self.depth_ptr
        if self.cheb_order:
            a.d_cheb_coef = self.aux['_cheb_coef'].data_ptr()
            a.d_cheb_x = self.aux['_cheb_x'].data_ptr()
        a.d_sens, a.d_dwl, a.d_dur_ms = (self.d_sens.data_ptr(), self.d_dwl.data_ptr(),
                                         self.d_dur.data_ptr())
        a.d_expected = d_expected.data_ptr() if d_expected is not None else None
        a.d_counts = d_counts.data_ptr() if d_counts is not None else None
        a.d_totals = self.d_totals.data_ptr()
        return a

    # ------------------------------------------------------------------
    def counts(self, mode, key=(0, 0), counts=None, want_expected=False):
        """Stage 1c: expected electrons and their integer draw.

        mode COUNT_NONE takes ``counts`` [N][W] from the caller (compat /
        deterministic); COUNT_ROUND / COUNT_POISSON compute them on the device."""
        e = self.e
        if want_expected and self.d_expected is None:
            self.d_expected = e.empty((self.N, self.W))
        if mode == _lib.COUNT_NONE:
            if counts is None:
                raise ValueError("counts required")
            c = np.ascontiguousarray(counts, dtype=np.int32)
            if c.shape != (self.N, self.W):
                raise ValueError("counts must be [N][W]")
            self.d_counts = e.to_dev(c)
        else:
            self.d_counts = e.empty((self.N, self.W), torch.int32)
        e.mark('k_counts', True)
        a = self._counts_args(mode, key, self.d_expected if want_expected else None, self.d_counts)
        check(lib.wb200_counts_ex(C.byref(a), e.stream_ptr()), "wb200_counts")
        e.mark('k_counts', False)
        self._release_depth()

    def expected_host(self):
        if self.d_expected is None:
            if getattr(self, '_depth_gone', False):
                raise RuntimeError("the planet-signal buffer of this exposure has been recycled; "
                                   "call counts(..., want_expected=True) to keep the expected counts")
            self.d_expected = self.e.empty((self.N, self.W))
            a = self._counts_args(_lib.COUNT_NONE, (0, 0), self.d_expected, None)
            check(lib.wb200_counts_ex(C.byref(a), self.e.stream_ptr()), "wb200_counts")
        return self.d_expected.cpu().numpy()

    # ------------------------------------------------------------------
    def _window_geometry(self, zmax, stride=1, margin_sigmas=4.0):
        """Per-sub-sample HBM windows that are guaranteed to contain every
        electron that lands inside the frame: trace extent +- zmax*sigma_max.
        ``stride`` > 1 evaluates every stride-th sub-sample only (enough for the
        bins-per-CTA choice of the direct path, which has no windows)."""
        g = self.grism
        from .grism import wavelength_calibration_coeffs, ANGSTROM_TO_MICRON
        x, y = self.xr_host[::stride], self.yr_host[::stride]
        m_t, c_t, m_w, c_w = wavelength_calibration_coeffs(x, y, g.trace_coeff, g.wl_solution)
        X0, X1 = x + 10, x + 20
        Y0 = m_t * (X0 - x) + c_t + y
        Y1 = m_t * (X1 - x) + c_t + y
        l0 = (m_w * np.sqrt((Y0 - y) ** 2 + (X0 - x) ** 2) + c_w) * ANGSTROM_TO_MICRON
        l1 = (m_w * np.sqrt((Y1 - y) ** 2 + (X1 - x) ** 2) + c_w) * ANGSTROM_TO_MICRON
        m_wl = (l1 - l0) / (X1 - X0)
        c_wl = l0 - m_wl * X0
        wl_lo, wl_hi = self.wl_host.min(), self.wl_host.max()
        xa = (wl_lo - c_wl) / m_wl
        xb = (wl_hi - c_wl) / m_wl
        ya = m_t * (xa - x) + c_t + y
        yb = m_t * (xb - x) + c_t + y
        x_lo = np.minimum(xa, xb) - self.sub_scale
        x_hi = np.maximum(xa, xb) - self.sub_scale
        y_lo = np.minimum(ya, yb) - self.sub_scale
        y_hi = np.maximum(ya, yb) - self.sub_scale
        sig = max(float(np.abs(np.polyval(self._poly[8:12], self.wl_host)).max()),
                  float(np.abs(np.polyval(self._poly[4:8], self.wl_host)).max()))
        margin = int(math.ceil(zmax * sig)) + 2
        L = self.L
        ok = np.isfinite(x_lo) & np.isfinite(x_hi) & np.isfinite(y_lo) & np.isfinite(y_hi)
        x_lo = np.where(ok, x_lo, 0.0)
        x_hi = np.where(ok, x_hi, 0.0)
        y_lo = np.where(ok, y_lo, 0.0)
        y_hi = np.where(ok, y_hi, 0.0)
        ox = np.clip(np.floor(x_lo).astype(np.int64) - margin, 0, L - 1)
        oy = np.clip(np.floor(y_lo).astype(np.int64) - margin, 0, L - 1)
        ex = np.clip(np.floor(x_hi).astype(np.int64) + margin + 1, 1, L)
        ey = np.clip(np.floor(y_hi).astype(np.int64) + margin + 1, 1, L)
        ww = int(max(1, (ex - ox).max()))
        wh = int(max(1, (ey - oy).max()))
        # chunk of bins whose trace segment plus 4 sigma fits the shared tile
        extent = float(np.abs(x_hi - x_lo).max()) + 1.0
        core = max(8.0, TILE_W - 2.0 * margin_sigmas * sig)
        chunk = int(self.W * core / extent) if extent > core else self.W
        chunk = max(32, chunk // 32 * 32)
        # equal chunks: the same number of CTAs per sub-sample, but of equal length (a short last
        # chunk makes CTAs of very different duration)
        n_chunks = int(math.ceil(self.W / float(chunk)))
        chunk = max(32, int(math.ceil(self.W / float(n_chunks) / 32.0)) * 32)
        want_ctas = 148 * 16
        if self.N * math.ceil(self.W / chunk) < want_ctas:
            per = max(1, want_ctas // self.N)
            chunk = min(chunk, max(32, int(math.ceil(self.W / per / 32.0)) * 32))
        return ox.astype(np.int32), oy.astype(np.int32), ww, wh, chunk

    def _gather_args(self, add_flat):
        """Geometry + flat planes shared by wb200_gather_flat and the direct thrower."""
        g = self.grism
        ga = _lib.GatherArgs()
        ga.n_reads = self.R
        ga.L, ga.F, ga.border = self.L, self.F, BORDER
        ga.add_flat = 1 if add_flat else 0
        ga.flat_off = g.flat_offset(self.S)
        ga.flat_n = 1014
        keep = None
        if add_flat:
            fl = g._load_flat()
            keep = self.e.cached_plane(("flat", g.name, g.flat_file_name), lambda: tuple(
                self.e.to_dev(p, np.float64) for p in fl['f']))
            ga.flat_n = int(fl['f'][0].shape[0])
            ga.flat_f32 = 1 if fl['f'][0].dtype.itemsize == 4 else 0
            ga.flat_wmin, ga.flat_wmax = float(fl['wmin']), float(fl['wmax'])
            for i in range(4):
                ga.d_flat[i] = keep[i].data_ptr()
        ga.d_read_end = self.d_read_end.data_ptr()
        return ga, keep

    def throw_direct(self, key=(0, 0), add_flat=True):
        """Native mode, stages 2+3+3b in one launch: Philox electrons binned in
        shared-memory tiles, flat applied at tile flush, accumulated into int64
        fixed-point interval planes (no per-sub-sample windows, no gather pass)."""
        e = self.e
        if self.d_counts is None:
            raise RuntimeError("counts() must run before throw_direct()")
        N, W, L, F = self.N, self.W, self.L, self.F
        self.d_acc = e.zeros((self.R, F, F), torch.int64)
        self.acc_fixed = True
        # the direct path has no windows: only the bins-per-CTA choice is needed, and it depends on
        # the trace length and the PSF width, not on this exposure's pointing -- cached per set-up
        gkey = ("chunk", self.grism.name, self.S, W, N, float(self.wl_host[0]), float(self.wl_host[-1]))
        # (3 sigma of the wide Gaussian either side of the chunk's trace segment inside the tile: the
        # few electrons of the end bins that leave it are replayed; fewer, longer CTAs pay the tile
        # zeroing / placement / flush less often -- measured: 2048 bins per CTA 1.91 ms, 1376 1.97,
        # 1024 2.01, 512 2.24 on the configs[3] shape)
        geo = e.cached_plane(gkey, lambda: self._window_geometry(ZMAX[_lib.RNG_PHILOX], stride=max(1, N // 32),
                                                                 margin_sigmas=3.0)[2:])
        ww, wh, chunk = geo
        self.win_geometry = (ww, wh, chunk)
        pa = _lib.PhotonArgs()
        pa.n_samples, pa.n_bins, pa.chunk_bins = N, W, chunk
        pa.nr, pa.nc = L, L
        pa.rng_mode = _lib.RNG_PHILOX
        pa.sub_scale = self.sub_scale
        pa.key0, pa.key1 = key[0] & 0xffffffff, key[1] & 0xffffffff
        pa.d_counts = self.d_counts.data_ptr()
        pa.d_totals = self.d_totals.data_ptr()
        pa.d_trace = self.d_trace.data_ptr()
        pa.d_wl = self.d_wl.data_ptr()
        pa.d_ratio, pa.d_sigl, pa.d_sigh = (self.d_ratio.data_ptr(), self.d_sigl.data_ptr(),
                                            self.d_sigh.data_ptr())
        ga, keep = self._gather_args(add_flat)
        ga.n_samples, ga.sample0 = N, 0
        ga.d_trace = self.d_trace.data_ptr()
        ga.d_acc = self.d_acc.data_ptr()
        self.tally = e.zeros((2,), torch.int64)
        pa.d_tally = self.tally.data_ptr()
        e.mark('k_throw', True)
        check(lib.wb200_throw_photons_direct(C.byref(pa), C.byref(ga), 0, e.stream_ptr()),
              "wb200_throw_photons_direct")
        e.mark('k_throw', False)
        self._keep = (keep,)
        return self.d_acc

    def throw(self, rng_mode, key=(0, 0), seeds=None, threads=1, normals=None, add_flat=True,
              window_cap=WINDOW_BYTES_CAP):
        """Stages 2+3 (+3b): throw and bin every electron into per-sub-sample HBM
        windows, then flat-field and accumulate IN SUB-SAMPLE ORDER into the
        float64 per-read-interval planes ``acc`` (the parity path; native mode
        uses throw_direct)."""
        e = self.e
        st = e.stream_ptr()
        N, W, L, F = self.N, self.W, self.L, self.F
        self.acc_fixed = False
        if self.d_counts is None:
            raise RuntimeError("counts() must run before throw()")
        d_off = d_seeds = d_norm = d_nbase = None
        if rng_mode == _lib.RNG_HOST:
            normals = np.ascontiguousarray(normals, dtype=np.float64)
            zmax = float(np.abs(normals[np.isfinite(normals)]).max()) * (1 + 1e-12) + 1e-9 \
                if normals.size else 1.0
        else:
            zmax = ZMAX[rng_mode]
        if rng_mode != _lib.RNG_PHILOX:
            d_off = e.empty((N, W), torch.int32)
            check(lib.wb200_count_offsets(N, W, _ptr(self.d_counts), _ptr(d_off), st),
                  "wb200_count_offsets")
        if rng_mode == _lib.RNG_RANDR:
            d_seeds = e.to_dev(seeds, np.int32)
        if rng_mode == _lib.RNG_HOST:
            totals = self.d_totals.cpu().numpy()
            base = np.zeros(N, dtype=np.int64)
            base[1:] = np.cumsum(2 * totals[:-1])
            if normals.size != int(2 * totals.sum()):
                raise ValueError("normals must hold 2*sum(counts) values (x then y per sub-sample)")
            d_norm = e.to_dev(normals)
            d_nbase = e.to_dev(base)

        ox, oy, ww, wh, chunk = self._window_geometry(zmax)
        self.win_geometry = (ww, wh, chunk)
        d_ox, d_oy = e.to_dev_many([ox, oy])
        per = ww * wh * 4
        nb = int(max(1, min(N, window_cap // per)))
        d_win = e.empty((nb, wh, ww), torch.int32)
        if self.d_acc is None:
            self.d_acc = e.zeros((self.R, F, F))
        else:
            self.d_acc.zero_()
        self.lost.zero_()

        g = self.grism
        fl = g._load_flat() if add_flat else None
        if add_flat:
            d_flat = e.cached_plane(("flat", g.name, g.flat_file_name), lambda: tuple(
                e.to_dev(p, np.float64) for p in fl['f']))
        for s0 in range(0, N, nb):
            n = min(nb, N - s0)
            d_win[:n].zero_()
            pa = _lib.PhotonArgs()
            pa.n_samples, pa.n_bins, pa.chunk_bins = n, W, chunk
            pa.nr, pa.nc = L, L
            pa.rng_mode, pa.threads = rng_mode, int(threads)
            pa.win_w, pa.win_h = ww, wh
            pa.sub_scale = self.sub_scale
            pa.key0, pa.key1 = key[0] & 0xffffffff, key[1] & 0xffffffff
            pa.d_counts = self.d_counts.data_ptr() + 4 * s0 * W
            pa.d_offsets = (d_off.data_ptr() + 4 * s0 * W) if d_off is not None else None
            pa.d_totals = self.d_totals.data_ptr() + 8 * s0
            pa.d_xpos = pa.d_ypos = None
            pa.d_trace = self.d_trace.data_ptr() + 8 * _lib.TRACE_STRIDE * s0
            pa.d_wl = self.d_wl.data_ptr()
            pa.d_ratio, pa.d_sigl, pa.d_sigh = (self.d_ratio.data_ptr(), self.d_sigl.data_ptr(),
                                                self.d_sigh.data_ptr())
            pa.d_seeds = (d_seeds.data_ptr() + 4 * s0) if d_seeds is not None else None
            pa.d_normals = d_norm.data_ptr() if d_norm is not None else None
            pa.d_normals_base = (d_nbase.data_ptr() + 8 * s0) if d_nbase is not None else None
            pa.d_win = d_win.data_ptr()
            pa.d_win_ox = d_ox.data_ptr() + 4 * s0
            pa.d_win_oy = d_oy.data_ptr() + 4 * s0
            pa.d_lost = self.lost.data_ptr()
            e.mark('k_throw', True)
            check(lib.wb200_throw_photons_at(C.byref(pa), s0, st), "wb200_throw_photons")
            e.mark('k_throw', False)

            ga = _lib.GatherArgs()
            ga.n_samples, ga.sample0, ga.n_reads = n, s0, self.R
            ga.L, ga.F, ga.border = L, F, BORDER
            ga.win_w, ga.win_h = ww, wh
            ga.add_flat = 1 if add_flat else 0
            ga.exact = 0 if rng_mode == _lib.RNG_PHILOX else 1
            ga.flat_off = g.flat_offset(self.S)
            if add_flat:
                ga.flat_n = int(fl['f'][0].shape[0])
                ga.flat_f32 = 1 if fl['f'][0].dtype.itemsize == 4 else 0
                ga.flat_wmin, ga.flat_wmax = float(fl['wmin']), float(fl['wmax'])
                for i in range(4):
                    ga.d_flat[i] = d_flat[i].data_ptr()
            else:
                ga.flat_n = 1014
            ga.d_read_end = self.d_read_end.data_ptr()
            ga.d_win = d_win.data_ptr()
            ga.d_win_ox = d_ox.data_ptr() + 4 * s0
            ga.d_win_oy = d_oy.data_ptr() + 4 * s0
            ga.d_trace = self.d_trace.data_ptr() + 8 * _lib.TRACE_STRIDE * s0
            ga.d_acc = self.d_acc.data_ptr()
            e.mark('k_gather', True)
            check(lib.wb200_gather_flat(C.byref(ga), st), "wb200_gather_flat")
            e.mark('k_gather', False)
        self._keep = (d_off, d_seeds, d_norm, d_nbase, d_ox, d_oy, d_win)
        return self.d_acc

    def check_lost(self):
        n = int(self.lost.item())
        if n:
            raise _lib.WayneB200Error(
                "{} electrons fell outside their sub-sample window (window bound violated)".format(n))

    def photons(self):
        """Electrons thrown in this exposure (incl. those landing off-frame)."""
        return int(self.d_totals.sum().item())

    def thrown(self):
        """The same as a 0-d device tensor (no synchronisation)."""
        return self.d_totals.sum()

    def counts_host(self):
        return self.d_counts.cpu().numpy()

    # ------------------------------------------------------------------
    def reads(self, dt_s, key=(0, 0), sky_rate=0.0, sky_plane=None, gain_plane=None, zero=None,
              dark=None, nl_planes=None, noise=(0.0, 0.0), clip=None, read_noise=0.0,
              cosmics=None, draws=None, exact_newton=False, out_f32=False, const_gain=2.35,
              sky_f32=True, fast_math=False):
        """Stage 4.  Planes are device tensors [F][F] (bordered) or None.

        dark = (dark[R][F][F], err[R][F][F]) device tensors; cosmics =
        (pixel[int32, bordered flat index], read[int32], energy[float64]) host arrays;
        draws = dict of host arrays for the compat mode: 'noise' [R][L][L],
        'sky' [R][L][L], 'dark' [R][F][F], 'rn' [R+1][F][F]."""
        e = self.e
        st = e.stream_ptr()
        F, R = self.F, self.R
        draws = draws or {}
        a = _lib.ReadsArgs()
        a.n_reads, a.F, a.border, a.out_f32 = R, F, BORDER, 1 if out_f32 else 0
        a.add_noise = 1 if ('noise' in draws or (noise[0] and noise[1])) else 0
        a.add_sky = 1 if ('sky' in draws or (sky_rate and sky_plane is not None)) else 0
        a.add_dark = 1 if ('dark' in draws or dark is not None) else 0
        a.add_nonlinear = 1 if nl_planes is not None else 0
        a.clip = 1 if clip is not None else 0
        a.add_read_noise = 1 if ('rn' in draws or read_noise) else 0
        a.exact_newton = 1 if exact_newton else 0
        a.key0, a.key1 = key[0] & 0xffffffff, key[1] & 0xffffffff
        a.noise_mean, a.noise_std = float(noise[0] or 0.0), float(noise[1] or 0.0)
        a.sky_rate = float(sky_rate or 0.0)
        a.sky_f32 = 1 if sky_f32 else 0
        a.fast_math = 1 if fast_math else 0
        a.acc_fixed = 1 if getattr(self, 'acc_fixed', False) else 0
        a.const_gain = float(const_gain)
        a.clip_lo, a.clip_hi = (float(clip[0]), float(clip[1])) if clip is not None else (0.0, 0.0)
        a.read_noise = float(read_noise or 0.0)
        keep = []
        d_dt = dt_s if isinstance(dt_s, torch.Tensor) else e.to_dev(dt_s, np.float64)
        a.d_dt = d_dt.data_ptr()
        a.d_acc = self.d_acc.data_ptr()
        a.d_sky = sky_plane.data_ptr() if sky_plane is not None else None
        a.d_gain = gain_plane.data_ptr() if gain_plane is not None else None
        a.d_zero = zero.data_ptr() if zero is not None else None
        if dark is not None:
            a.d_dark, a.d_dark_err = dark[0].data_ptr(), dark[1].data_ptr()
        if nl_planes is not None:
            for i in range(7):
                a.d_nl[i] = nl_planes[i].data_ptr()

        def pad_stack(arr, n):
            arr = np.asarray(arr, dtype=np.float64)
            if arr.shape[-1] == F:
                return np.ascontiguousarray(arr)
            out = np.zeros((n, F, F))
            m = arr.shape[-1]
            out[:, BORDER:BORDER + m, BORDER:BORDER + m] = arr
            return out

        for name, field, n in (('noise', 'd_draw_noise', R), ('sky', 'd_draw_sky', R),
                               ('dark', 'd_draw_dark', R), ('rn', 'd_draw_rn', R + 1)):
            if name in draws:
                t = e.to_dev(pad_stack(draws[name], n))
                keep.append(t)
                setattr(a, field, t.data_ptr())
        if cosmics is not None and len(cosmics[0]):
            pix, rd, en = cosmics
            if isinstance(pix, torch.Tensor):
                d_pix, d_rd, d_en = pix, rd, en
            else:
                d_pix, d_rd, d_en = e.to_dev_many([np.asarray(pix, dtype=np.int32),
                                                   np.asarray(rd, dtype=np.int32),
                                                   np.asarray(en, dtype=np.float64)])
            d_head = e.empty((F * F,), torch.int32)
            d_next = e.empty((len(pix),), torch.int32)
            check(lib.wb200_cosmic_chains(len(pix), _ptr(d_pix), F * F, _ptr(d_head), _ptr(d_next),
                                          st), "wb200_cosmic_chains")
            a.n_cosmics = len(pix)
            a.d_cos_head, a.d_cos_next = d_head.data_ptr(), d_next.data_ptr()
            a.d_cos_read, a.d_cos_energy = d_rd.data_ptr(), d_en.data_ptr()
            keep += [d_pix, d_rd, d_en, d_head, d_next]
        d_iters = e.zeros((16,), torch.int32)
        a.d_newton_iters = d_iters.data_ptr()
        out = e.empty((R + 1, F, F), torch.float32 if out_f32 else torch.float64)
        a.d_out = out.data_ptr()
        e.mark('k_reads', True)
        check(lib.wb200_reads(C.byref(a), st), "wb200_reads")
        e.mark('k_reads', False)
        self._keep_reads = keep + [d_dt, d_iters]
        self.newton_iters = d_iters
        return out


# ---------------------------------------------------------------------------------------
# Exposure-level interface: one wb200_exposure_run call per exposure
# ---------------------------------------------------------------------------------------
def _f32_plane(a):
    """Host plane in the calibration file's float32, native byte order, contiguous.
    Every value must survive the round trip (the library checks float64 input itself;
    this keeps the upload at half the bytes)."""
    a = np.asarray(a)
    out = np.ascontiguousarray(a, dtype=np.float32)
    if a.dtype.itemsize > 4 and not np.array_equal(out.astype(np.float64), np.asarray(a, dtype=np.float64),
                                                   equal_nan=True):
        raise ValueError("calibration plane is not float32-representable")
    return out


class ContextRun(object):
    """What an exposure queued through the context leaves behind for the caller (the
    counterpart of ExposureRun for tests, the bench and ExposureGenerator.photons)."""

    def __init__(self, ctx, N, W, R, stats):
        self.ctx, self.N, self.W, self.R = ctx, N, W, R
        self.stats = stats                 # device int64 [4]: thrown, binned, dropped, 0
        self.tally = stats[1:3]
        self.lost = None
        self.win_geometry = (None, None, ctx.info()[0])

    def thrown(self):
        return self.stats[0]

    def photons(self):
        """Electrons thrown (incl. those landing off-frame).  When the exposure's reads were
        fetched to the host the statistics travelled with them: no device synchronisation."""
        host = getattr(self, 'stats_host', None)
        if host is not None:
            self.stats_done.synchronize()
            return int(host[0])
        return int(self.stats[0].item())

    def counts_host(self):
        return self.ctx.read_scratch(0, (self.N, self.W), np.int32)

    def totals_host(self):
        return self.ctx.read_scratch(1, (self.N,), np.uint64)

    def trace_host(self):
        return self.ctx.read_scratch(2, (self.N, _lib.TRACE_STRIDE), np.float64)

    def tables_host(self):
        return tuple(self.ctx.read_scratch(3, (5, self.W), np.float64))


class ExposureContext(object):
    """Python face of a ``wb200_ctx``: the resident planes of one instrument configuration
    (grism, SUBARRAY, SAMPSEQ) on one GPU and the single call that queues an exposure."""

    def __init__(self, engine, grism, detector, subarray, sampseq):
        self.e, self.grism, self.det = engine, grism, detector
        self.S, self.seq = int(subarray), sampseq
        self.L = 1014 if self.S == 1024 else self.S
        self.F = min(self.S + 10, 1024)
        self._h = C.c_void_p()
        check(lib.wb200_ctx_create(engine.device.index, C.byref(self._h)), "wb200_ctx_create")
        self._have = set()
        self._dark_reads = 0
        self._seen_inputs = {}
        self._set_instrument()

    def __del__(self):
        try:
            if self._h:
                lib.wb200_ctx_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:      # interpreter shutdown
            pass

    def _check(self, rc, what):
        if rc != _lib.OK:
            msg = lib.wb200_ctx_last_error(self._h)
            raise _lib.WayneB200Error("{} failed ({}): {}".format(
                what, rc, msg.decode("utf-8", "replace") if msg else ""))

    def _set_instrument(self):
        g, det = self.grism, self.det
        sens_wl, sens_val = g._load_sens()
        fl_n, f32, wmin, wmax = 1014, 1, 0.0, 1.0
        inst = _lib.Instrument()
        inst.subarray, inst.L, inst.F, inst.border = self.S, self.L, self.F, BORDER
        inst.flat_off = g.flat_offset(self.S)
        inst.flat_n, inst.flat_f32, inst.flat_wmin, inst.flat_wmax = fl_n, f32, wmin, wmax
        inst.n_sens = len(sens_wl)
        inst.sub_scale = float(507 - self.S // 2)
        for i, v in enumerate(np.concatenate([np.asarray(p.coeffs, dtype=np.float64) for p in
                                              (g.psf_ratio_poly, g.psf_sigmal_poly, g.psf_sigmah_poly)])):
            inst.psf_poly12[i] = v
        for i in range(9):
            inst.trace_coeff9[i] = float(g.trace_coeff[i])
            inst.wl_sol9[i] = float(g.wl_solution[i])
        inst.const_gain = float(det.constant_gain)
        inst.clip_lo, inst.clip_hi = float(det.min_counts), float(det.max_counts)
        inst.read_noise = float(det.read_noise)
        self._inst = inst
        self._check(lib.wb200_ctx_set_instrument(self._h, C.byref(inst)), "wb200_ctx_set_instrument")
        self._upload(_lib.PLANE_SENS_WL, np.ascontiguousarray(sens_wl, dtype=np.float64), _lib.F64)
        self._upload(_lib.PLANE_SENS_VAL, np.ascontiguousarray(sens_val, dtype=np.float64), _lib.F64)

    keep_host_copies = False     # tests: remember what was uploaded / passed (write_bundle)

    def _upload(self, which, arr, dtype):
        self._check(lib.wb200_ctx_upload_plane(self._h, which, C.c_void_p(arr.ctypes.data), dtype, arr.size),
                    "wb200_ctx_upload_plane")
        self._have.add(which)
        if self.keep_host_copies:
            self.__dict__.setdefault('_uploaded', {})[which] = (arr, dtype)

    def write_bundle(self, path):
        """Everything a C host needs to repeat the LAST exposure of this context, as the flat record
        file examples/c_host/exposure_host.c reads (needs keep_host_copies = True beforehand)."""
        import struct
        a, host = self._last_args
        tags = dict(wl=110, flux=111, xref=112, yref=113, dur_ms=114, dt_s=115, read_end=116, sep_row=117,
                    sep_col=118, cos_pixel=119, cos_read=120, cos_energy=121)
        with open(path, 'wb') as f:
            def rec(tag, dtype, arr):
                arr = np.ascontiguousarray(arr)
                f.write(struct.pack('<iiq', tag, dtype, arr.size))
                f.write(arr.tobytes())
            for which, (arr, dtype) in sorted(self._uploaded.items()):
                rec(which, 1 if dtype == _lib.F64 else 0, arr)
            rec(100, 3, np.frombuffer(bytes(self._inst), dtype=np.uint8))
            rec(101, 3, np.frombuffer(bytes(a), dtype=np.uint8))
            for name, arr in host.items():
                rec(tags[name], {np.dtype(np.float64): 1, np.dtype(np.int32): 2}[arr.dtype], arr)

    def ensure_planes(self, add_flat, sky, gain, nonlinear, dark_reads, zero_read):
        """Upload (once) the planes the requested terms read.  dark_reads = number of reads
        that need dark planes (0: none); zero_read = host array or None."""
        g, det, e = self.grism, self.det, self.e
        if add_flat and _lib.PLANE_FLAT0 not in self._have:
            fl = g._load_flat()
            inst = self._inst
            inst.flat_n = int(fl['f'][0].shape[0])
            inst.flat_f32 = 1 if fl['f'][0].dtype.itemsize == 4 else 0
            inst.flat_wmin, inst.flat_wmax = float(fl['wmin']), float(fl['wmax'])
            self._check(lib.wb200_ctx_set_instrument(self._h, C.byref(inst)), "wb200_ctx_set_instrument")
            for i in range(4):
                self._upload(_lib.PLANE_FLAT0 + i, _f32_plane(fl['f'][i]), _lib.F32)
        if sky and _lib.PLANE_SKY not in self._have:
            self._upload(_lib.PLANE_SKY, _f32_plane(e.bordered(g.get_master_sky(self.L), self.F)), _lib.F32)
        if gain and _lib.PLANE_GAIN not in self._have:
            self._upload(_lib.PLANE_GAIN, _f32_plane(e.bordered(det.get_gain(self.S), self.F, fill=1.0)), _lib.F32)
        if nonlinear and _lib.PLANE_NL0 not in self._have:
            for i, p in enumerate(det.non_linear_planes(self.F)[:4]):
                self._upload(_lib.PLANE_NL0 + i, _f32_plane(p), _lib.F32)
        if dark_reads > self._dark_reads:
            planes = [det.dark_planes(r + 1, self.S, self.seq) for r in range(1, dark_reads + 1)]
            self._upload(_lib.PLANE_DARK, _f32_plane(np.stack([p[0] for p in planes])), _lib.F32)
            self._upload(_lib.PLANE_DARK_ERR, _f32_plane(np.stack([p[1] for p in planes])), _lib.F32)
            self._dark_reads = dark_reads
        if zero_read is not None and _lib.PLANE_ZERO not in self._have:
            self._upload(_lib.PLANE_ZERO, np.ascontiguousarray(zero_read, dtype=np.float64), _lib.F64)

    STAGES = ('k_stage1', 'k_counts', 'k_cosmic_chains', 'k_throw', 'k_reads')

    def profile(self, on):
        self._check(lib.wb200_ctx_profile(self._h, 1 if on else 0), "wb200_ctx_profile")

    def stage_times(self):
        """{stage: (total ms, launches)} since the last call (synchronises the device)."""
        ms, n = (C.c_double * 8)(), (_lib.c_i64 * 8)()
        self._check(lib.wb200_ctx_stage_times(self._h, ms, n), "wb200_ctx_stage_times")
        return {name: (ms[i], n[i]) for i, name in enumerate(self.STAGES) if n[i]}

    def info(self):
        out = (_lib.c_i64 * 8)()
        self._check(lib.wb200_ctx_info(self._h, out), "wb200_ctx_info")
        return list(out)

    def read_scratch(self, which, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        self._check(lib.wb200_ctx_read_scratch(self._h, which, C.c_void_p(out.ctypes.data), out.nbytes),
                    "wb200_ctx_read_scratch")
        return out

    def _device_input_ready(self, t):
        """True when the device tensor ``t`` is known to be complete, i.e. nothing still queued on
        the current stream produces it.  The first time a tensor (same storage, same version
        counter) is seen, an event is recorded on the current stream; once that event has
        completed the tensor is ready for good.  Lets the library start the next exposure's
        tables and counts on its own stream while the previous exposure is still running
        (wb200_exposure_args.device_inputs_ready) -- a visit keeps its stellar flux and its
        planet-signal coefficients resident and hands the same tensors to every exposure.  (A
        tensor rewritten in place by something torch does not see -- a foreign kernel through
        data_ptr() -- must be handed over as a new tensor object.)"""
        # keyed by the tensor OBJECT (a weak reference keeps the entry honest): a new tensor that the
        # allocator put at a recycled address is a different object and is not ready
        seen = self._seen_inputs
        ent = seen.get(id(t))
        if ent is not None and ent[0]() is t and ent[1] == t._version and ent[2] == t.data_ptr():
            return ent[3].query()
        if len(seen) > 64:
            seen.clear()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.e.device))
        seen[id(t)] = (weakref.ref(t), t._version, t.data_ptr(), ev)
        return False

    def run(self, wl_um, flux, depth, depth_col0, xr, yr, dur_ms, dt_s, read_end, scale, key, count_mode,
            add_flat, sky_rate, add_gain, add_dark, add_nonlinear, clip, add_read_noise, add_zero, noise,
            cosmics, out_f32):
        """Queue one exposure on the current stream; returns (reads tensor [R+1][F][F], ContextRun)."""
        e = self.e
        dev_inputs = []    # device-resident inputs of this exposure (see _device_input_ready)
        W, N, R = len(wl_um), len(xr), len(read_end)
        a = _lib.ExposureArgs()
        a.n_samples, a.n_bins, a.n_reads, a.count_mode = N, W, R, count_mode
        a.add_flat, a.add_sky, a.add_gain = int(bool(add_flat)), int(bool(sky_rate)), int(bool(add_gain))
        a.add_dark, a.add_nonlinear, a.clip = int(bool(add_dark)), int(bool(add_nonlinear)), int(bool(clip))
        a.add_read_noise, a.add_zero = int(bool(add_read_noise)), int(bool(add_zero))
        a.add_noise = int(bool(noise[0] and noise[1]))
        a.out_f32 = int(bool(out_f32))
        a.key0, a.key1 = key[0] & 0xffffffff, key[1] & 0xffffffff
        a.scale = 1.0 if scale is None else float(scale)
        a.sky_rate = float(sky_rate or 0.0)
        a.noise_mean, a.noise_std = float(noise[0] or 0.0), float(noise[1] or 0.0)
        keep = []          # host arrays must stay alive until the call has copied them (it does before returning)

        named = {}

        def host(arr, dtype=np.float64, name=None):
            arr = np.ascontiguousarray(arr, dtype=dtype)
            keep.append(arr)
            if name:
                named[name] = arr
            return C.c_void_p(arr.ctypes.data)

        a.wl, a.xref, a.yref = host(wl_um, name='wl'), host(xr, name='xref'), host(yr, name='yref')
        a.dur_ms, a.dt_s = host(dur_ms, name='dur_ms'), host(dt_s, name='dt_s')
        a.read_end = host(read_end, np.int32, name='read_end')
        if isinstance(flux, torch.Tensor):
            keep.append(flux)
            a.d_flux = flux.data_ptr()
            dev_inputs.append(flux)
        else:
            a.flux = host(flux, name='flux')
        slot = None
        if depth is not None:
            if hasattr(depth, 'row') and hasattr(depth, 'col'):          # lightcurve.SeparableSignal
                a.sep_row = host(depth.row[:N], name='sep_row')
                a.sep_col = host(depth.col[depth_col0:depth_col0 + W], name='sep_col')
            elif hasattr(depth, 'coef') and hasattr(depth, 'x'):          # lightcurve.ChebyshevSignal
                a.cheb_order = int(depth.coef.shape[1])
                a.cheb_x = host(depth.x[depth_col0:depth_col0 + W])
                coef = depth.coef[:N]
                if isinstance(coef, torch.Tensor):
                    if not coef.is_contiguous():
                        coef = coef.contiguous()
                        dev_inputs.append(None)        # made just now on the current stream: not ready
                    else:
                        dev_inputs.append(coef)
                    keep.append(coef)
                    a.d_cheb_coef = coef.data_ptr()
                else:
                    a.cheb_coef = host(coef)
            else:
                if isinstance(depth, torch.Tensor):
                    if depth.dtype != torch.float64:
                        raise ValueError("device planet_signal must be float64")
                    d_depth = e.to_dev(depth)
                    dev_inputs.append(d_depth if d_depth is depth else None)
                else:
                    depth = np.asarray(depth)
                    if depth.dtype != np.float64 or not depth.flags.c_contiguous:
                        depth = np.ascontiguousarray(depth, dtype=np.float64)
                    # host array: the library uploads it on its own upload stream, behind this
                    # exposure's small arrays (a pinned array makes that copy asynchronous)
                    keep.append(depth)
                    a.depth_ld = int(depth.shape[1])
                    a.depth = C.c_void_p(depth.ctypes.data + 8 * int(depth_col0))
                    d_depth = None
                if d_depth is not None:
                    keep.append(d_depth)
                    a.depth_ld = int(depth.shape[1])
                    a.d_depth = d_depth.data_ptr() + 8 * int(depth_col0)
        if cosmics is not None and len(cosmics[0]):
            a.n_cosmics = len(cosmics[0])
            a.cos_pixel = host(cosmics[0], np.int32, name='cos_pixel')
            a.cos_read = host(cosmics[1], np.int32, name='cos_read')
            a.cos_energy = host(cosmics[2], name='cos_energy')
        stats = e.empty((4,), torch.int64)
        a.d_stats = stats.data_ptr()
        a.device_inputs_ready = int(all(t is not None and self._device_input_ready(t) for t in dev_inputs))
        out = e.empty((R + 1, self.F, self.F), torch.float32 if out_f32 else torch.float64)
        self.profile(e.profile)
        e.mark('exposure', True)
        self._check(lib.wb200_exposure_run(self._h, C.byref(a), C.c_void_p(out.data_ptr()), e.stream_ptr()),
                    "wb200_exposure_run")
        e.mark('exposure', False)
        if slot is not None:
            e.release_upload(slot)
        if self.keep_host_copies:
            self._last_args = (a, named)
        return out, ContextRun(self, N, W, R, stats)
