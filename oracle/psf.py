"""TEST INFRASTRUCTURE ONLY -- ctypes access to the electron-thrower oracles.

  oracle (port)      oracle/psf_oracle.c, our plain-C restatement of
                     wayne/pyparallel_menu.c:10-113
  reference          oracle/_ref/libwayne_ref_psf.so = the reference's
                     pyparallel_menu.c compiled UNMODIFIED by oracle/Makefile
                     (present when built in a container that has /root/reference;
                     the built file travels to the GPU box, the sources do not)
  reference wrapper  oracle/_ref/pyparallel*.so = the reference's Cython module
                     (pyparallel.pyx) built unmodified

Never imported by wayne_b200/.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "_build", "libpsf_oracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libwayne_ref_psf.so")
REF_PYX = os.path.join(HERE, "_ref", "pyparallel" + sysconfig.get_config_var("EXT_SUFFIX"))

DP = C.POINTER(C.c_double)
IP = C.POINTER(C.c_int)


def build(quiet=True):
    """Build the C restatement (and the reference .so when /root/reference exists)."""
    res = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    if not quiet:
        print(res.stdout)


_port = None
_ref = None
_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.free.restype = None
_libc.rand_r.argtypes = [C.POINTER(C.c_uint)]
_libc.rand_r.restype = C.c_int


def port():
    global _port
    if _port is None:
        if not os.path.isfile(PORT_LIB):
            build()
        lib = C.CDLL(PORT_LIB)
        lib.wo_rand_r.argtypes = [C.POINTER(C.c_uint32)]
        lib.wo_rand_r.restype = C.c_int
        lib.wo_fill_normals.argtypes = [DP, C.c_int, C.c_int, C.c_int]
        lib.wo_fill_normals.restype = None
        lib.wo_bin_electrons.argtypes = [IP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, DP,
                                         C.c_long, IP]
        lib.wo_bin_electrons.restype = C.c_long
        lib.wo_psf.argtypes = [IP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_int, C.c_int,
                               IP]
        lib.wo_psf.restype = C.c_int
        _port = lib
    return _port


def have_reference():
    return os.path.isfile(REF_LIB)


def reference():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_LIB)
        lib.PSF.argtypes = [IP, C.c_int, DP, DP, DP, DP, DP, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.PSF.restype = IP
        _ref = lib
    return _ref


def _prep(counts, x, y, ratio, sigl, sigh):
    c = np.ascontiguousarray(counts, dtype=np.int32)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, ratio, sigl, sigh)]
    return c, arrs


def libc_rand_r(seed, n):
    s = C.c_uint(seed)
    return [_libc.rand_r(C.byref(s)) for _ in range(n)]


def port_rand_r(seed, n):
    s = C.c_uint32(seed)
    return [port().wo_rand_r(C.byref(s)) for _ in range(n)]


def fill_normals(ssum, test, threads):
    """The reference's A[2*ssum] table for (test, threads) (pyparallel_menu.c:40-64)."""
    A = np.empty(2 * int(ssum) + 1, dtype=np.float64)
    port().wo_fill_normals(A.ctypes.data_as(DP), int(ssum), int(test), int(threads))
    return A[:2 * int(ssum)]


def bin_electrons(counts, x, y, ratio, sigl, sigh, nr, nc, normals):
    """Scatter loop only (pyparallel_menu.c:87-108) with a supplied A table."""
    c, (x, y, r, sl, sh) = _prep(counts, x, y, ratio, sigl, sigh)
    A = np.ascontiguousarray(normals, dtype=np.float64)
    ssum = int(c.sum())
    assert A.size == 2 * ssum
    frame = np.zeros(nr * nc, dtype=np.int32)
    port().wo_bin_electrons(c.ctypes.data_as(IP), len(c), x.ctypes.data_as(DP), y.ctypes.data_as(DP),
                            r.ctypes.data_as(DP), sl.ctypes.data_as(DP), sh.ctypes.data_as(DP),
                            nr, nc, A.ctypes.data_as(DP), ssum, frame.ctypes.data_as(IP))
    return frame.reshape(nr, nc)


def psf_port(counts, x, y, ratio, sigl, sigh, nr, nc, test, threads):
    """== PSF() by the C restatement; int32 [nr][nc]."""
    c, (x, y, r, sl, sh) = _prep(counts, x, y, ratio, sigl, sigh)
    frame = np.zeros(nr * nc, dtype=np.int32)
    rc = port().wo_psf(c.ctypes.data_as(IP), len(c), x.ctypes.data_as(DP), y.ctypes.data_as(DP),
                       r.ctypes.data_as(DP), sl.ctypes.data_as(DP), sh.ctypes.data_as(DP), nr, nc,
                       int(test), int(threads), frame.ctypes.data_as(IP))
    if rc:
        raise MemoryError("wo_psf")
    return frame.reshape(nr, nc)


def psf_reference(counts, x, y, ratio, sigl, sigh, nr, nc, test, threads):
    """The UNMODIFIED reference PSF(); int32 [nr][nc]."""
    c, (x, y, r, sl, sh) = _prep(counts, x, y, ratio, sigl, sigh)
    p = reference().PSF(c.ctypes.data_as(IP), len(c), x.ctypes.data_as(DP), y.ctypes.data_as(DP),
                        r.ctypes.data_as(DP), sl.ctypes.data_as(DP), sh.ctypes.data_as(DP), nr, nc,
                        int(test), int(threads))
    try:
        return np.ctypeslib.as_array(p, shape=(nr * nc,)).copy().reshape(nr, nc)
    finally:
        _libc.free(p)


def reference_pyparallel():
    """The reference's Cython module (apply_psf), or None when it was not built."""
    if not os.path.isfile(REF_PYX):
        return None
    spec = importlib.util.spec_from_file_location("pyparallel", REF_PYX)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def psf_case(seed=0, n_bins=512, mean_count=40.0, frame=256, x0=40.0, x1=200.0, y0=120.0):
    """A seeded PSF() input set shaped like one G141 sub-sample."""
    rng = np.random.default_rng(seed)
    wl = np.linspace(0.988, 1.777, n_bins)
    counts = rng.poisson(mean_count, n_bins).astype(np.int32)
    x = np.linspace(x0, x1, n_bins) + rng.normal(0, 1e-3, n_bins)
    y = y0 + 0.009 * (x - x0) + rng.normal(0, 1e-3, n_bins)
    ratio = np.polyval([-0.25063428, 0.8332488, -0.80546074, 0.39896516], wl)
    sigl = np.polyval([0.69245668, -2.1043046, 2.22284446, -0.29689335], wl)
    sigh = np.polyval([2.90366189, -8.81859432, 8.96049229, 2.254503], wl)
    return dict(counts=counts, x=x, y=y, ratio=ratio, sigl=sigl, sigh=sigh, nr=frame, nc=frame)
