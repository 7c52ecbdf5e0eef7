"""Import alias: `import smplk` loads the package directory `3d-human-body-reconstruction_b200/`
(whose name is not a valid Python identifier) under the module name `smplk`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "3d-human-body-reconstruction_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
