"""Import shim: the package directory is named ``yoxel-voxel_b200`` (not a Python identifier),
so ``import yoxel_voxel_b200`` resolves its sub-modules from that directory."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__.insert(0, _os.path.join(_os.path.dirname(_here), "yoxel-voxel_b200"))

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
