"""Cosmic-ray hit generators (mirror of wayne/trend_generators/cosmic_rays.py).

A generator answers three questions -- how many hits in ``time`` seconds, with
what energies, and where -- and ``cosmic_frame`` composes them.  Besides the
dense-frame API of the reference, ``cosmic_hits`` returns the sparse hit list
(rows, cols, energies) that the CUDA per-pixel pass consumes; it draws from the
numpy global RandomState in the SAME order as ``cosmic_frame`` (count, energies,
rows, cols), so compat-mode exposures consume the reference's random stream.
"""
import numpy as np


class BaseCosmicGenerator(object):
    """11 hits per second of energy 25000, whatever the array size."""

    full_frame_rate = 11

    def __init__(self):
        pass

    def _number_of_cosmics(self, time, size=1024):
        return self.full_frame_rate * time

    def _rate_full_frame_to_size(self, full_frame_rate, size):
        npix = size * size if isinstance(size, (int, np.integer)) else size[0] * size[1]
        return full_frame_rate / (1024. * 1024.) * npix

    def _generate_cosmic_energies(self, number):
        return [25000] * number

    def _generate_array(self, size):
        try:
            n = int(size)
            return np.zeros((n, n))
        except TypeError:
            return np.zeros(size)

    def _cosmics_to_array(self, list_of_energies, array):
        n = len(list_of_energies)
        rows = np.random.randint(0, len(array), n)
        cols = np.random.randint(0, len(array[0]), n)
        for i, energy in enumerate(list_of_energies):
            array[rows[i], cols[i]] += energy
        return array

    def cosmic_hits(self, time, size=1024):
        """(rows, cols, energies) of the hits in ``time`` seconds."""
        number = self._number_of_cosmics(time, size)
        energies = self._generate_cosmic_energies(number)
        shape = self._generate_array(size).shape
        n = len(energies)
        rows = np.random.randint(0, shape[0], n)
        cols = np.random.randint(0, shape[1], n)
        return rows, cols, np.asarray(energies, dtype=np.float64).reshape(n)

    def cosmic_frame(self, time, size=1024):
        number = self._number_of_cosmics(time, size)
        energies = self._generate_cosmic_energies(number)
        return self._cosmics_to_array(energies, self._generate_array(size))


class MinMaxPossionCosmicGenerator(BaseCosmicGenerator):
    """Poisson number of hits at ``rate`` per second per 1024x1024 pixels, uniform
    integer energies in [min_count, max_count)."""

    def __init__(self, rate=11., min_count=10000, max_count=35000):
        BaseCosmicGenerator.__init__(self)
        self.rate = rate
        self.min_count = min_count
        self.max_count = max_count

    def _number_of_cosmics(self, time, size=1024):
        return np.random.poisson(self._rate_full_frame_to_size(self.rate, size) * time)

    def _generate_cosmic_energies(self, number):
        energies = np.random.randint(self.min_count, self.max_count, number)
        if number == 1:
            energies = [energies]
        return energies
