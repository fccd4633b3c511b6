"""Import alias: the package directory is named `motion-diffusion-model_b200` (not a Python identifier), so
`import b200mdm` resolves to it through this shim."""
import importlib
import sys

_pkg = importlib.import_module("motion-diffusion-model_b200")
sys.modules[__name__] = _pkg
