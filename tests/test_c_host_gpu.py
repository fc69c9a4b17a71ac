"""A plain-C host of the library (examples/c_host/exposure_host.c: gcc, no Python, no torch)
runs an exposure through wb200_ctx_* / wb200_exposure_run and gets, bit for bit, the reads the
Python host layer gets for the same inputs -- the boundary really is a C ABI."""
import os
import subprocess

import numpy as np
import pytest

from tests import harness

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_host_reproduces_the_python_hosts_exposure(calb_dir, tmp_path):
    from wayne import detector, grism
    from wayne import units as u
    from wayne.exposure_generator import ExposureGenerator
    from wayne_b200 import lightcurve as lc
    from wayne_b200.engine import DeviceEngine, ExposureContext
    exe = str(tmp_path / 'exposure_host')
    cuda = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-O2', '-I', os.path.join(ROOT, 'include'),
                    '-I', os.path.join(cuda, 'include'), os.path.join(ROOT, 'examples', 'c_host', 'exposure_host.c'),
                    '-L', os.path.join(ROOT, 'wayne_b200'), '-lwayne_b200', '-L', os.path.join(cuda, 'lib64'),
                    '-lcudart', '-Wl,-rpath,' + os.path.join(ROOT, 'wayne_b200'), '-o', exe], check=True)
    DeviceEngine.get().drop_planes()                     # a fresh context, recording what it is given
    ExposureContext.keep_host_copies = True
    try:
        wl, flux, planet = harness.spectrum(level=3.0e-14)
        eg = ExposureGenerator(detector.WFC3_IR(), grism.G141(), 5, 'SPARS10', 256, None, rng='philox')
        _, mid, dur, ri = eg._gen_scanning_sample_times(200 * u.ms)
        n = len(np.asarray(u.value_in(mid, u.ms)))
        sig = lc.SeparableSignal(0.5 * (1 + np.tanh(np.linspace(-2, 2, n))), planet)
        exp = eg.scanning_frame(404.5, 457.4, 0.02, 0.02, wl * u.micron, flux, sig, 7.4325 * u.pixel / u.s,
                                200 * u.ms, mid, dur, ri, cosmic_rate=11., sky_background=5.5 * u.count / u.s,
                                scale_factor=0.9991, rng_key=(1963, 42))
        want = np.array([r[0] for r in exp.reads])
        ctx = DeviceEngine.get().exposure_context(eg.grism, eg.detector, 256, 'SPARS10')
        bundle, out = str(tmp_path / 'bundle.bin'), str(tmp_path / 'reads.bin')
        ctx.write_bundle(bundle)
    finally:
        ExposureContext.keep_host_copies = False
        DeviceEngine.get().drop_planes()
    res = subprocess.run([exe, bundle, out], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    got = np.fromfile(out, dtype=np.float64).reshape(want.shape)
    assert want.shape == (5, 266, 266) and np.abs(want[-1]).max() > 100
    assert np.array_equal(got, want)
    assert 'electrons thrown %d,' % eg.photons in res.stdout and 'dropped' in res.stdout
