"""ctypes binding of libwayne_b200.so (include/wayne_b200.h).

The library is the ONLY compute path of this package: if it is missing or fails
to load, importing this module raises -- there is no CPU fallback (the CPU
restatement lives in oracle/ and is test infrastructure only).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WAYNE_B200_LIB") or os.path.join(HERE, "libwayne_b200.so")

OK = 0
RNG_PHILOX, RNG_RANDR, RNG_HOST = 0, 1, 2
COUNT_NONE, COUNT_ROUND, COUNT_POISSON = 0, 1, 2
TRACE_STRIDE = 8

c_void_p = C.c_void_p
c_int = C.c_int
c_i32 = C.c_int32
c_i64 = C.c_int64
c_u32 = C.c_uint32
c_u64 = C.c_uint64
c_double = C.c_double
DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


class PhotonArgs(C.Structure):
    _fields_ = [
        ("n_samples", c_i32), ("n_bins", c_i32), ("chunk_bins", c_i32),
        ("nr", c_i32), ("nc", c_i32), ("rng_mode", c_i32), ("threads", c_i32),
        ("win_w", c_i32), ("win_h", c_i32),
        ("sub_scale", c_double),
        ("key0", c_u32), ("key1", c_u32),
        ("d_counts", c_void_p), ("d_offsets", c_void_p), ("d_totals", c_void_p),
        ("d_xpos", c_void_p), ("d_ypos", c_void_p), ("d_trace", c_void_p),
        ("d_wl", c_void_p), ("d_ratio", c_void_p), ("d_sigl", c_void_p), ("d_sigh", c_void_p),
        ("d_seeds", c_void_p), ("d_normals", c_void_p), ("d_normals_base", c_void_p),
        ("d_win", c_void_p), ("d_win_ox", c_void_p), ("d_win_oy", c_void_p),
        ("d_lost", c_void_p), ("d_tally", c_void_p),
    ]


class CountsArgs(C.Structure):
    _fields_ = [
        ("n_samples", c_i32), ("n_bins", c_i32), ("count_mode", c_i32), ("cheb_order", c_i32),
        ("key0", c_u32), ("key1", c_u32),
        ("scale", c_double), ("depth_ld", c_i64),
        ("d_flux", c_void_p), ("d_depth", c_void_p), ("d_cheb_coef", c_void_p), ("d_cheb_x", c_void_p),
        ("d_sens", c_void_p), ("d_dwl", c_void_p), ("d_dur_ms", c_void_p),
        ("d_expected", c_void_p), ("d_counts", c_void_p), ("d_totals", c_void_p),
        ("d_sep_row", c_void_p),
    ]


class GatherArgs(C.Structure):
    _fields_ = [
        ("n_samples", c_i32), ("sample0", c_i32), ("n_reads", c_i32),
        ("L", c_i32), ("F", c_i32), ("border", c_i32),
        ("win_w", c_i32), ("win_h", c_i32),
        ("add_flat", c_i32), ("flat_off", c_i32), ("flat_n", c_i32),
        ("flat_f32", c_i32), ("exact", c_i32), ("flat_planes_f32", c_i32),
        ("flat_wmin", c_double), ("flat_wmax", c_double),
        ("d_read_end", c_void_p), ("d_win", c_void_p), ("d_win_ox", c_void_p),
        ("d_win_oy", c_void_p), ("d_trace", c_void_p),
        ("d_flat", c_void_p * 4),
        ("d_acc", c_void_p),
    ]


class ReadsArgs(C.Structure):
    _fields_ = [
        ("n_reads", c_i32), ("F", c_i32), ("border", c_i32), ("out_f32", c_i32),
        ("add_noise", c_i32), ("add_sky", c_i32), ("add_dark", c_i32),
        ("add_nonlinear", c_i32), ("clip", c_i32), ("add_read_noise", c_i32),
        ("exact_newton", c_i32), ("n_cosmics", c_i32),
        ("key0", c_u32), ("key1", c_u32),
        ("noise_mean", c_double), ("noise_std", c_double), ("sky_rate", c_double),
        ("sky_f32", c_i32), ("fast_math", c_i32), ("acc_fixed", c_i32), ("zero_acc", c_i32),
        ("const_gain", c_double), ("clip_lo", c_double), ("clip_hi", c_double),
        ("read_noise", c_double),
        ("d_dt", c_void_p), ("d_acc", c_void_p), ("d_sky", c_void_p), ("d_gain", c_void_p),
        ("d_zero", c_void_p), ("d_dark", c_void_p), ("d_dark_err", c_void_p),
        ("d_nl", c_void_p * 7),
        ("d_draw_noise", c_void_p), ("d_draw_sky", c_void_p), ("d_draw_dark", c_void_p),
        ("d_draw_rn", c_void_p),
        ("d_cos_head", c_void_p), ("d_cos_next", c_void_p), ("d_cos_read", c_void_p),
        ("d_cos_energy", c_void_p),
        ("d_newton_iters", c_void_p),
        ("d_out", c_void_p),
        ("planes_f32", c_i32), ("pad2", c_i32),
    ]


class Instrument(C.Structure):
    _fields_ = [
        ("subarray", c_i32), ("L", c_i32), ("F", c_i32), ("border", c_i32),
        ("flat_off", c_i32), ("flat_n", c_i32), ("flat_f32", c_i32), ("n_sens", c_i32),
        ("sub_scale", c_double), ("flat_wmin", c_double), ("flat_wmax", c_double),
        ("psf_poly12", c_double * 12), ("trace_coeff9", c_double * 9), ("wl_sol9", c_double * 9),
        ("const_gain", c_double), ("clip_lo", c_double), ("clip_hi", c_double), ("read_noise", c_double),
    ]


class ExposureArgs(C.Structure):
    _fields_ = [
        ("n_samples", c_i32), ("n_bins", c_i32), ("n_reads", c_i32),
        ("count_mode", c_i32), ("cheb_order", c_i32), ("n_cosmics", c_i32),
        ("add_flat", c_i32), ("add_sky", c_i32), ("add_gain", c_i32), ("add_dark", c_i32),
        ("add_nonlinear", c_i32), ("clip", c_i32), ("add_read_noise", c_i32), ("add_zero", c_i32),
        ("add_noise", c_i32), ("out_f32", c_i32),
        ("key0", c_u32), ("key1", c_u32), ("device_inputs_ready", c_i32),
        ("scale", c_double), ("sky_rate", c_double), ("noise_mean", c_double), ("noise_std", c_double),
        ("depth_ld", c_i64),
        ("wl", c_void_p), ("flux", c_void_p), ("xref", c_void_p), ("yref", c_void_p), ("dur_ms", c_void_p),
        ("dt_s", c_void_p), ("read_end", c_void_p), ("cheb_x", c_void_p), ("cheb_coef", c_void_p),
        ("sep_row", c_void_p), ("sep_col", c_void_p), ("depth", c_void_p),
        ("cos_pixel", c_void_p), ("cos_read", c_void_p), ("cos_energy", c_void_p),
        ("d_depth", c_void_p), ("d_cheb_coef", c_void_p), ("d_flux", c_void_p), ("d_stats", c_void_p),
    ]


(PLANE_FLAT0, PLANE_FLAT1, PLANE_FLAT2, PLANE_FLAT3, PLANE_SKY, PLANE_GAIN, PLANE_NL0, PLANE_NL1, PLANE_NL2,
 PLANE_NL3, PLANE_DARK, PLANE_DARK_ERR, PLANE_ZERO, PLANE_SENS_WL, PLANE_SENS_VAL) = range(15)
F32, F64 = 0, 1

# every symbol include/wayne_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "wb200_last_error": (C.c_char_p, []),
    "wb200_version": (c_int, []),
    "wb200_device_count": (c_int, []),
    "wb200_launch_count": (c_u64, []),
    "PSF": (IP, [IP, c_int, DP, DP, DP, DP, DP, c_int, c_int, c_int, c_int]),
    "wb200_psf_host": (c_int, [IP, c_int, DP, DP, DP, DP, DP, c_int, c_int, c_int, c_int, c_int,
                               DP, IP]),
    "wb200_bin_tables": (c_int, [c_int, c_void_p, DP, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "wb200_trace_table": (c_int, [c_int, c_void_p, c_void_p, DP, DP, c_void_p, c_void_p]),
    "wb200_trace_positions": (c_int, [c_int, c_int, c_void_p, c_void_p, c_double, c_void_p,
                                      c_void_p, c_void_p]),
    "wb200_counts": (c_int, [c_int, c_int, c_void_p, c_void_p, c_i64, c_void_p, c_void_p,
                             c_void_p, c_double, c_int, c_u32, c_u32, c_void_p, c_void_p,
                             c_void_p, c_void_p]),
    "wb200_counts_ex": (c_int, [C.POINTER(CountsArgs), c_void_p]),
    "wb200_transit_cheb": (c_int, [c_int, c_int, c_void_p, c_double, c_double, DP, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "wb200_count_offsets": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "wb200_throw_photons": (c_int, [C.POINTER(PhotonArgs), c_void_p]),
    "wb200_throw_photons_at": (c_int, [C.POINTER(PhotonArgs), c_int, c_void_p]),
    "wb200_gather_flat": (c_int, [C.POINTER(GatherArgs), c_void_p]),
    "wb200_throw_photons_direct": (c_int, [C.POINTER(PhotonArgs), C.POINTER(GatherArgs), c_int, c_void_p]),
    "wb200_reads": (c_int, [C.POINTER(ReadsArgs), c_void_p]),
    "wb200_cosmic_chains": (c_int, [c_int, c_void_p, c_i32, c_void_p, c_void_p, c_void_p]),
    "wb200_ctx_create": (c_int, [c_int, C.POINTER(c_void_p)]),
    "wb200_ctx_destroy": (c_int, [c_void_p]),
    "wb200_ctx_last_error": (C.c_char_p, [c_void_p]),
    "wb200_ctx_set_instrument": (c_int, [c_void_p, C.POINTER(Instrument)]),
    "wb200_ctx_upload_plane": (c_int, [c_void_p, c_int, c_void_p, c_int, c_i64]),
    "wb200_exposure_run": (c_int, [c_void_p, C.POINTER(ExposureArgs), c_void_p, c_void_p]),
    "wb200_ctx_info": (c_int, [c_void_p, C.POINTER(c_i64)]),
    "wb200_ctx_profile": (c_int, [c_void_p, c_int]),
    "wb200_ctx_stage_times": (c_int, [c_void_p, DP, C.POINTER(c_i64)]),
    "wb200_ctx_read_scratch": (c_int, [c_void_p, c_int, c_void_p, c_i64]),
    "wb200_microbench": (c_int, [c_int, c_int, DP, DP]),
    "wb200_philox_words": (c_int, [c_int, C.POINTER(c_u32), C.POINTER(c_u32), C.POINTER(c_u32)]),
}


class WayneB200Error(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libwayne_b200.so is not built ({}); run `python -m wayne_b200.build` -- there is no "
            "CPU fallback for the exposure path".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc, what=""):
    if rc != OK:
        msg = lib.wb200_last_error()
        raise WayneB200Error("{} failed ({}): {}".format(
            what or "libwayne_b200 call", rc, msg.decode("utf-8", "replace") if msg else ""))


def launch_count():
    return int(lib.wb200_launch_count())
