"""Visit-long trends: one flux scale factor per exposure (mirror of
wayne/trend_generators/visit_trends.py).  The factor reaches the path as the
``scale_factor`` argument of scanning_frame (exposure_generator.py:620-621)."""
import abc

import numpy as np

from .. import units as u


class BaseVisitTrend(object):
    __metaclass__ = abc.ABCMeta

    def __init__(self, visit_plan, coeffs=None):
        self.visit_plan = visit_plan
        self.coeffs = coeffs
        self.scale_factors = self._gen_scaling_factors(visit_plan, coeffs)

    def _gen_scaling_factors(self, visit_plan, coeffs):
        raise NotImplementedError

    def get_scale_factor(self, exp_num):
        return self.scale_factors[exp_num]


def _days(t):
    if u.is_quantity(t):
        return np.asarray(t.to(u.day).value, dtype=float)
    return np.array([u.value_in(v, u.day) for v in t], dtype=float) \
        if np.ndim(t) and len(t) and u.is_quantity(t[0]) else np.asarray(t, dtype=float)


class HookAndLongTermRamp(BaseVisitTrend):
    """(1 - a1 (t - t_v)) (1 - b1 exp(-b2 (t - t_orbit)))."""

    def _gen_scaling_factors(self, visit_plan, coeffs):
        t = _days(visit_plan['exp_start_times'])
        t_0 = gen_orbit_start_times_per_exp(t, visit_plan['orbit_start_index'])
        return self.ramp_model(t, t_0, *coeffs)

    @staticmethod
    def ramp_model(t, t_0, a1, b1, b2, to):
        t = _days(t)
        return (1 - a1 * (t - to)) * (1 - b1 * np.exp(-b2 * (t - t_0)))


def gen_orbit_start_times_per_exp(time_array, obs_start_index):
    """For every exposure, the start time of the orbit it belongs to."""
    time_array = _days(time_array)
    starts = list(obs_start_index) + [len(time_array)]
    t_0 = np.zeros(len(time_array))
    for a, b in zip(starts[:-1], starts[1:]):
        t_0[a:b] = time_array[a]
    return t_0
