"""Import alias: ``import hvlm_b200`` == the package in ``handsonvlm-release_b200/`` (whose directory name,
fixed by the project layout, is not a valid Python identifier).  Submodules resolve to the SAME module objects under both
names (``hvlm_b200.tower is sys.modules['handsonvlm-release_b200.tower']``): one copy of every class and registry."""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_REAL = "handsonvlm-release_b200"
_pkg = importlib.import_module(_REAL)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``hvlm_b200.<sub>`` -> the already imported (or now imported) ``handsonvlm-release_b200.<sub>`` module."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(__name__ + "."):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        real = importlib.import_module(_REAL + spec.name[len(__name__):])
        self._specs[spec.name] = real.__spec__
        return real

    def exec_module(self, module):
        # the import machinery has just pointed module.__spec__ at the alias spec; the module keeps its real one (its
        # relative imports resolve against __package__ / __spec__.parent == the real package)
        for name, real_spec in list(self._specs.items()):
            if sys.modules.get(real_spec.name) is module:
                module.__spec__ = real_spec
                del self._specs[name]

    _specs: dict = {}


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[__name__] = _pkg
