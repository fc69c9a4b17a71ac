"""Timeline of a few pipelined exposures (torch.profiler / CUPTI; nsys is not in
the image): prints every kernel / memcpy with stream, start and duration, plus
the host time spent issuing each exposure.
    python tools/trace_pipeline.py [workload] [resident:0|1] [n]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

wk = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c4']
resident = (sys.argv[2] if len(sys.argv) > 2 else '0') == '1'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
bench.calibration_dir(wk)
inp = bench.make_inputs(wk)
from wayne import units as u  # noqa: E402
from wayne.exposure_generator import ExposureGenerator  # noqa: E402

N, W = len(inp['lightcurve']), len(inp['wl'])
depth_pin = torch.empty((N, W), dtype=torch.float64, pin_memory=True)
depth = depth_pin.numpy()
depth[:] = inp['depth0'][None, :] * inp['lightcurve'][:, None]
print('from_numpy(pinned view).is_pinned():', torch.from_numpy(depth).is_pinned())
depth_dev = depth_pin.cuda()
wl_q = inp['wl'] * u.micron


def one(i):
    eg = ExposureGenerator(*inp['eg_args'], rng='philox', device=0)
    kw = bench.frame_kwargs(wk, i)
    return eg.scanning_frame(kw.pop('x_ref'), kw.pop('y_ref'), kw.pop('x_jitter'), kw.pop('y_jitter'),
                             wl_q, inp['flux'], depth_dev if resident else depth, kw.pop('scan_speed'),
                             kw.pop('sample_rate'), inp['mid'], inp['dur'], inp['read_index'],
                             rng_key=(1963, i), device_result=resident, **kw)


pend = [one(i) for i in range(5)]
if not resident:
    for e in pend:
        e.reads
torch.cuda.synchronize()
del pend
host = []
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    t00 = time.perf_counter()
    pend = []
    for i in range(n):
        t0 = time.perf_counter()
        pend.append(one(10 + i))
        host.append((time.perf_counter() - t0) * 1e3)
        if not resident and len(pend) > 2:
            t0 = time.perf_counter()
            pend.pop(0).reads
            host[-1] = (host[-1], (time.perf_counter() - t0) * 1e3)
    if not resident:
        for e in pend:
            e.reads
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t00) * 1e3
print('wall ms per exposure: %.2f; host issue ms (issue, wait-for-reads): %s' % (wall / n, host))
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
busy = 0.0
for e in evs:
    d = e.time_range.end - e.time_range.start
    if d >= 30:          # >= 30 us
        print('%9.3f ms  +%8.3f ms  %s' % ((e.time_range.start - t0) / 1e3, d / 1e3, e.name[:70]))
