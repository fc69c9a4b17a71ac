"""cProfile of the host side of one exposure (bench workload), to find Python /
numpy overhead around the kernel launches.  Usage (GPU box):
    python tools/profile_host.py [workload] [resident:0|1]"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

wk = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c4']
resident = (sys.argv[2] if len(sys.argv) > 2 else '1') == '1'
import torch  # noqa: E402

bench.calibration_dir(wk)
inp = bench.make_inputs(wk)
from wayne import units as u  # noqa: E402
from wayne.exposure_generator import ExposureGenerator  # noqa: E402

depth = inp['depth0'][None, :] * inp['lightcurve'][:, None]
depth_dev = torch.from_numpy(depth).cuda()
wl_q = inp['wl'] * u.micron


def one(i):
    eg = ExposureGenerator(*inp['eg_args'], rng='philox', device=0)
    kw = bench.frame_kwargs(wk, i)
    return eg.scanning_frame(kw.pop('x_ref'), kw.pop('y_ref'), kw.pop('x_jitter'), kw.pop('y_jitter'),
                             wl_q, inp['flux'], depth_dev if resident else depth, kw.pop('scan_speed'),
                             kw.pop('sample_rate'), inp['mid'], inp['dur'], inp['read_index'],
                             rng_key=(1963, i), device_result=resident, **kw)


for i in range(3):
    one(i)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
N_PROF = int(sys.argv[3]) if len(sys.argv) > 3 else 5
for i in range(N_PROF):
    one(10 + i)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
pstats.Stats(pr).sort_stats('tottime').print_stats(30)
