"""Debug aid: wrap the engine's host-side entry points and report calls that take
longer than a threshold (to find sporadic host stalls).  Usage:
    WB200_HOSTTRACE=5 python bench.py ...      (threshold in ms; printed to stderr)"""
import functools
import os
import sys
import time


def install():
    thr = float(os.environ.get('WB200_HOSTTRACE', '0') or 0)
    if thr <= 0:
        return
    from wayne_b200 import engine, exposure_generator

    def wrap(owner, name):
        fn = getattr(owner, name)

        @functools.wraps(fn)
        def inner(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                dt = (time.perf_counter() - t0) * 1e3
                if dt >= thr:
                    sys.stderr.write('[hosttrace] t=%.1f %s.%s %.2f ms\n' % (time.perf_counter() * 1e3 % 1e6, owner.__name__, name, dt))
        setattr(owner, name, inner)

    for n in ('upload_async', 'to_dev_many', 'pinned_out', 'fetch_async', 'admit', 'retire', 'to_dev',
              'zeros', 'empty', 'mark'):
        wrap(engine.DeviceEngine, n)
    for n in ('__init__', 'counts', 'throw_direct', 'throw', 'reads', '_window_geometry', '_gather_args'):
        wrap(engine.ExposureRun, n)
    for n in ('scanning_frame', '_device_planes', '_gen_zero_read', '__init__', '_gen_scanning_sample_times'):
        wrap(exposure_generator.ExposureGenerator, n)
    import gc
    state = {}

    def on_gc(phase, info):
        if phase == 'start':
            state['t'] = time.perf_counter()
        else:
            dt = (time.perf_counter() - state.get('t', time.perf_counter())) * 1e3
            if dt >= thr:
                sys.stderr.write('[hosttrace] t=%.1f gc gen%d %.2f ms (%d collected)\n' % (
                    time.perf_counter() * 1e3 % 1e6, info.get('generation', -1), dt, info.get('collected', 0)))
    gc.callbacks.append(on_gc)
