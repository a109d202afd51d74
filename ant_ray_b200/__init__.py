"""Import shim: the package directory is named ``ant-ray_b200`` (not a Python identifier),
so ``import ant_ray_b200`` resolves here and re-exports that directory as this package."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "ant-ray_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
