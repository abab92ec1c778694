"""cvx_proj_b200 -- B200-native APAP (moving-DLT + mesh warp) hot path.

Drop-in for the reference's ``pyviz/apap.py`` + ``pyviz/apap_utils.py`` call surface:
``APAP(gamma, sigma, final_size, offset).local_homography / local_warp`` and
``get_mesh / get_vertice / final_size / uniform_blend``.  Python host code over a C-ABI
shared library of hand-written sm_100a CUDA kernels (``csrc/``, ``include/apap_b200.h``).
There is no CPU fallback: every compute entry point raises if the library or a GPU is
missing.
"""
from .apap_utils import final_size, get_mesh, get_vertice, uniform_blend

from .driver import mat_layout, save2mat, stitch_pair

__all__ = ["APAP", "final_size", "get_mesh", "get_vertice", "uniform_blend", "mat_layout", "save2mat", "stitch_pair"]
__version__ = "0.1.0"


def __getattr__(name):  # APAP pulls in torch; keep `import cvx_proj_b200.synth` light
    if name == "APAP":
        from .apap import APAP
        return APAP
    raise AttributeError(name)
