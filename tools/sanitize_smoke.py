"""Small run of every kernel for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wayne_b200 import calibration, params, pyparallel  # noqa: E402

d = os.path.join(tempfile.gettempdir(), 'wayne_b200_synth_calb')
calibration.write_synthetic_calibration(d, modes=((256, 'SPARS10'),))
params.set_calibration_dir(d)
from oracle import psf as O  # noqa: E402
from tests import harness  # noqa: E402
from wayne import detector, grism  # noqa: E402
from wayne import units as u  # noqa: E402
from wayne.exposure_generator import ExposureGenerator  # noqa: E402

case = O.psf_case(seed=1, n_bins=256, mean_count=20.0)
args = (case["counts"], case["x"], case["y"], case["ratio"], case["sigl"], case["sigh"], 256, 256)
for rng in ('randr', 'philox'):
    f = pyparallel.psf_frame(*args, test=3, threads=2, rng=rng)
    print(rng, int(f.sum()))
A = np.random.default_rng(1).standard_normal(2 * int(case["counts"].sum()))
print('host', int(pyparallel.psf_frame(*args, rng='host', normals=A).sum()))
wl, flux, planet = harness.spectrum(level=2.0e-15, n_wl=200)
for mode in ('numpy', 'philox'):
    for direct in (True, False):
        params.direct_accumulation = direct
        eg = ExposureGenerator(detector.WFC3_IR(), grism.G141(), 5, 'SPARS10', 256, None, rng=mode)
        np.random.seed(3)
        exp = eg.scanning_frame(404.5, 457.4, 0.02, 0.02, wl * u.micron, flux, None, 7.4 * u.pixel / u.s,
                                2000 * u.ms, cosmic_rate=11., noise_mean=0.5, noise_std=0.1, rng_key=(1, 2))
        print(mode, direct, eg.photons, float(exp.reads[-1][0].sum()))
