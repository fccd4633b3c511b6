"""Import alias: the package directory is named `motion-diffusion-model_b200` (not a Python identifier), so
`import b200mdm` -- and every `b200mdm.<submodule>` -- resolves to the very same module objects through this shim.

A plain `sys.modules['b200mdm'] = pkg` is not enough: `from b200mdm.utils.sampler_util import X` would then execute
the submodule a second time under the alias name and hand out a second class object (isinstance checks inside the
package would fail silently).  The finder below maps any `b200mdm.a.b` to the already-importable
`motion-diffusion-model_b200.a.b` and registers it under both names.
"""
import importlib
import importlib.abc
import importlib.util
import sys

_REAL = "motion-diffusion-model_b200"
_ALIAS = "b200mdm"


class _AliasLoader(importlib.abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return importlib.import_module(self.real_name)      # the one and only module object

    def exec_module(self, module):                           # already executed under its real name
        pass

    def get_code(self, fullname):                            # runpy (`python -m b200mdm.build`) asks for this
        spec = importlib.util.find_spec(self.real_name)
        return spec.loader.get_code(self.real_name)


class _AliasFinder(importlib.abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname != _ALIAS and not fullname.startswith(_ALIAS + "."):
            return None
        real = _REAL + fullname[len(_ALIAS):]
        try:
            real_spec = importlib.util.find_spec(real)
        except (ImportError, ValueError):
            return None
        if real_spec is None:
            return None
        spec = importlib.util.spec_from_loader(fullname, _AliasLoader(real), origin=real_spec.origin,
                                               is_package=real_spec.submodule_search_locations is not None)
        spec.has_location = real_spec.has_location
        return spec


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
_pkg = importlib.import_module(_REAL)
sys.modules[_ALIAS] = _pkg
