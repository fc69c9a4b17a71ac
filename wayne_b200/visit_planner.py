"""Visit planner: exposure start times of an HST visit (mirror of
wayne/visit_planner.py:5-129, same name, arguments and output dictionary)."""
import numpy as np

from . import units as u


def VisitPlanner(detector, NSAMP, SAMPSEQ, SUBARRAY, num_orbits=3, time_per_orbit=54 * u.min,
                 hst_period=95 * u.min, exp_overhead=1 * u.min):
    """Start time of every exposure [minutes from 0]: per orbit, guide-star
    acquisition (6 min first orbit, 5 min after), then exposures of
    ``exptime + exp_overhead`` until the target sets, with a 5.8 min buffer dump
    whenever more than ``num_exp_per_buffer`` exposures have accumulated."""
    exptime_min = float(u.value_in(detector.exptime(NSAMP, SUBARRAY, SAMPSEQ), u.min))
    exp_per_dump = detector.num_exp_per_buffer(NSAMP, SUBARRAY)
    per_orbit = float(u.value_in(time_per_orbit, u.min))
    period = float(u.value_in(hst_period, u.min))
    overhead = float(u.value_in(exp_overhead, u.min))
    buffer_dump = 5.8
    exp_times, orbit_start_index, buffer_dump_index = [], [], []
    for orbit in range(num_orbits):
        orbit_start_index.append(len(exp_times))
        start = period * orbit
        t = start + (6.0 if orbit == 0 else 5.0)
        end = start + per_orbit
        n = 0
        while t < end:
            exp_times.append(t)
            t += exptime_min + overhead
            n += 1
            if n > exp_per_dump:
                t += buffer_dump
                n = 0
                buffer_dump_index.append(len(exp_times))
    return {
        'exp_times': np.array(exp_times) * u.min,
        'NSAMP': NSAMP, 'SAMPSEQ': SAMPSEQ, 'SUBARRAY': SUBARRAY,
        'num_exp': len(exp_times),
        'exptime': detector.exptime(NSAMP, SUBARRAY, SAMPSEQ),
        'num_orbits': num_orbits, 'exp_overhead': exp_overhead,
        'time_per_orbit': time_per_orbit, 'hst_period': hst_period,
        'buffer_dump_index': buffer_dump_index, 'orbit_start_index': orbit_start_index,
    }
