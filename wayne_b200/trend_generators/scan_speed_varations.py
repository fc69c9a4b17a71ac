"""Scan-speed variations: rescale the duration of every sub-sample (mirror of
wayne/trend_generators/scan_speed_varations.py; module name keeps the
reference's spelling).  Host-side: the arrays have one entry per sub-sample."""
import numpy as np

from .. import units as u


class SSVSine(object):
    """duration *= 1 + stddev/100 * sin(period * (y - y_0) + phase)."""

    def __init__(self, stddev=1.5, period=0.7, start_phase='rand'):
        self.stddev = stddev
        self.period = period
        self.start_phase = start_phase

    def _scaling(self, y_mid_points, phase):
        dy = np.asarray(y_mid_points, dtype=float) - y_mid_points[0]
        return (self.stddev / 100.) * np.sin((self.period * dy) + phase) + 1.

    def get_subsample_exposure_times(self, y_mid_points, sample_durations,
                                     subsample_exptime=None, total_exptime=None):
        if self.start_phase == 'rand':
            # the reference raises AttributeError here; defined
            # behaviour: random phase, total exposure kept equal to phase 0
            phase = np.random.random() * 2 * np.pi
            ref_mean = np.mean(self._scaling(y_mid_points, 0.))
            scaling = self._scaling(y_mid_points, phase)
            scaling = scaling * (ref_mean / np.mean(scaling))
        else:
            scaling = self._scaling(y_mid_points, self.start_phase)
        return sample_durations * scaling


class SSVModulatedSine(object):
    """Stochastic amplitude/period-modulated sine with optional blips; total
    exposure time and each read time are preserved to the microsecond by moving
    1 us quanta between sub-samples.  Returns (durations [ms], read indexes)."""

    def __init__(self, amplitude=10, period=1.1, blip_proba=1):
        self.amplitude = amplitude
        self.period = period
        self.blip_proba = blip_proba

    def get_subsample_exposure_times(self, y_mid_points, sample_durations, read_times, sample_rate):
        rnd = np.random
        read_times = np.asarray(u.value_in(read_times, u.s), dtype=float)
        sample_rate = float(u.value_in(sample_rate, u.s))
        exptime = np.round(read_times[-1], 6)
        tt = np.arange(0, exptime, sample_rate)
        quantum = 0.000001

        def slow_sine():
            return rnd.normal(0.1, 0.05) * np.sin(
                (2 * np.pi / rnd.normal(2.0 * exptime, 0.5 * exptime)) * tt
                + rnd.random() * 2 * np.pi)

        amp = 1.0 + slow_sine()
        if 100.0 * rnd.random() < self.blip_proba:
            amp = amp + rnd.normal(1.0, 0.1) * np.exp(
                -(tt - rnd.random() * exptime) ** 2 / (2 * (self.period / 2) ** 2))
        final_amp = sample_rate * (self.amplitude / 100.0) * amp
        final_per = self.period * (1.0 + slow_sine())
        phase = rnd.random() * 2 * np.pi
        sub = np.round(sample_rate + final_amp * np.sin((2 * np.pi / final_per) * tt + phase), 6)

        def spread(diff, lo, hi, sign):
            for _ in range(abs(diff)):
                sub[rnd.randint(lo, hi)] += sign * quantum

        diff = int((10 ** 6) * np.round(exptime - np.sum(sub), 6))
        spread(diff, 0, len(sub), -1 if diff < 0 else 1)

        breaks = [int(np.argmin(abs(np.cumsum(sub) - t))) for t in read_times]

        diff = int((10 ** 6) * np.round(read_times[0] - np.sum(sub[:breaks[0] + 1]), 6))
        sign = -1 if diff < 0 else 1
        for i in np.int_(rnd.power(3, abs(diff)) * (breaks[0] + 1)):
            sub[breaks[0] - i] += sign * quantum
            sub[rnd.randint(breaks[0] + 1, len(sub))] -= sign * quantum
        for r in range(1, len(read_times) - 1):
            diff = int((10 ** 6) * np.round(read_times[r] - np.sum(sub[:breaks[r] + 1]), 6))
            sign = -1 if diff < 0 else 1
            for _ in range(abs(diff)):
                sub[rnd.randint(breaks[r - 1] + 1, breaks[r] + 1)] += sign * quantum
                sub[rnd.randint(breaks[r] + 1, len(sub))] -= sign * quantum
        return (sub * u.s).to(u.ms), breaks
