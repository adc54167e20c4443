"""Import shim: the package sources live in `hy-video-prfl_b200/` (the directory name the build
contract asks for, which is not a valid Python identifier); `import prfl_b200` resolves there."""
import os as _os

__path__.append(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "hy-video-prfl_b200"))
from ._version import __version__  # noqa: E402,F401
