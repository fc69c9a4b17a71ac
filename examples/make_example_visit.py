"""Writes a runnable example visit modelled on the reference's shipped example
(examples/hd209458b_12181_simulation_parameters.yml: HD 209458 b, G141 spatial
scan, SUBARRAY 256, NSAMP 5 SPARS10, 5 orbits, 121 exposures) with SYNTHETIC
input tables (planet spectrum, pointing, sky, exposure times) and a black-body
star in place of the PHOENIX spectrum the reference cannot ship either.

    python examples/make_example_visit.py [outdir]          # default: examples/hd209458b_like
    python -m wayne_b200.run_visit -p examples/hd209458b_like/params.yml

The calibration set must exist (WAYNE_CALB_DIR); a synthetic one is written with
    python -c "from wayne_b200 import calibration; calibration.write_synthetic_calibration('calb')"
"""
import os
import sys

import numpy as np
import yaml

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                         'hd209458b_like')
os.makedirs(out, exist_ok=True)
rng = np.random.default_rng(12181)

# planet spectrum: transit depth vs wavelength, R ~ 11000 over 0.9-1.8 micron (water band at 1.4)
wl = 0.9 * np.exp(np.arange(7600) / 11000.0)
wl = wl[wl < 1.8]
depth = 0.01463 + 1.6e-4 * np.exp(-0.5 * ((wl - 1.4) / 0.06) ** 2) + 4e-5 * np.exp(-0.5 * ((wl - 1.15) / 0.04) ** 2)
np.savetxt(os.path.join(out, 'planet_spectrum.dat'), np.c_[wl, depth], fmt='%.8f')

# 5 orbits of 95 min, ~24 exposures each: 22.3 s exposure + ~80 s overhead
t0 = 2456196.28836
starts = []
for orbit in range(5):
    t = t0 - 0.16 + orbit * 95.0 / 1440.0
    n = 25 if orbit else 21
    starts += list(t + np.arange(n) * (22.317 + 82.0) / 86400.0)
starts = np.array(starts)
np.savetxt(os.path.join(out, 'jd.txt'), starts, fmt='%.8f')
n = len(starts)
np.savetxt(os.path.join(out, 'xref.txt'), 404.0 + np.cumsum(0.004 * rng.standard_normal(n)), fmt='%.5f')
np.savetxt(os.path.join(out, 'yref.txt'), 457.3 + np.cumsum(0.002 * rng.standard_normal(n)), fmt='%.5f')
np.savetxt(os.path.join(out, 'sky.txt'), 5.5 + 0.8 * np.sin(np.linspace(0, 5 * 2 * np.pi, n)), fmt='%.4f')

cfg = {
    'general': {'oec_location': False, 'outdir': 'simulated', 'seed': 1963, 'threads': 4},
    'target': {'name': 'HD 209458 b', 'planet_spectrum_file': 'planet_spectrum.dat', 'rebin_resolution': False,
               'stellar_spectrum_file': False, 'stellar_temperature': 6065, 'flux_scale': 2.0e-19,
               'period': 3.524746, 'sma': 0.047309, 'stellar_radius': 1.155, 'inclination': 86.71,
               'eccentricity': 0.0, 'periastron': 0.0, 'transit_time': t0,
               'ldcoeffs': [0.800627, -0.757066, 0.897268, -0.384804]},
    'observation': {'detector': 'WFC3IR', 'grism': 'G141', 'x_ref': 'xref.txt', 'y_ref': 'yref.txt',
                    'NSAMP': 5, 'SAMPSEQ': 'SPARS10', 'SUBARRAY': 256, 'start_JD': False,
                    'exp_start_times': 'jd.txt', 'num_orbits': 5, 'sample_rate': 10, 'spatial_scan': True,
                    'scan_speed': 7.4325, 'ssv_type': 'sine', 'ssv_coeffs': [1.5, 1.1, 0], 'x_shifts': 0,
                    'x_jitter': 0.025, 'y_shifts': 0, 'y_jitter': 1e-15, 'noise_mean': False,
                    'noise_std': False, 'add_dark': True, 'add_flat': True, 'add_gain_variations': True,
                    'add_non_linear': True, 'add_read_noise': True, 'add_initial_bias': True,
                    'add_stellar_noise': True, 'sky_background': 'sky.txt', 'cosmic_rate': 11,
                    'clip_values_det_limits': True},
    'trends': {'visit_trend_coeffs': [0.005, 0.0011, 400, t0]},
}
with open(os.path.join(out, 'params.yml'), 'w') as f:
    yaml.safe_dump(cfg, f, sort_keys=False)
print('wrote', out, '(%d exposures, %d spectral elements)' % (n, len(wl)))
