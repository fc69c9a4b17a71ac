"""``wayne`` -- import-compatible alias of :mod:`wayne_b200`.

Code written against the reference (``from wayne import grism, detector``,
``import wayne.exposure_generator``, ``wayne.pyparallel.apply_psf`` ...) resolves
to the B200-native modules of the same names; nothing is re-implemented here.
"""
import importlib
import importlib.abc
import importlib.util
import sys

_REAL = "wayne_b200"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(__name__ + "."):
            return None
        real = _REAL + fullname[len(__name__):]
        try:
            if importlib.util.find_spec(real) is None:
                return None
        except ModuleNotFoundError:
            return None
        return importlib.util.spec_from_loader(fullname, self)

    _real_specs = {}

    def create_module(self, spec):
        mod = importlib.import_module(_REAL + spec.name[len(__name__):])
        self._real_specs[spec.name] = mod.__spec__
        return mod

    def exec_module(self, module):
        # the import machinery has just pointed __spec__ at the alias; the
        # module keeps its own identity (relative imports inside it rely on it)
        for alias, real in self._real_specs.items():
            if real is not None and real.name == module.__name__:
                module.__spec__ = real


sys.meta_path.insert(0, _AliasFinder())


def __getattr__(name):
    try:
        return importlib.import_module(__name__ + "." + name)
    except ImportError:
        raise AttributeError(name)
