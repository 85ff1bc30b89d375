"""Import shim: the package directory is `camels-diffusion-model_b200/` (a hyphen is
not a valid Python identifier), so this module exposes it as the importable
package `camels_diffusion_model_b200`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "camels-diffusion-model_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _fh:
    exec(compile(_fh.read(), __file__, "exec"))
