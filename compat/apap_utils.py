"""Drop-in module name of the reference's ``pyviz/apap_utils.py``: put this directory on ``sys.path``
(where the reference's ``pyviz/`` was) and ``from apap_utils import *`` (pyviz/apap.py:15) resolves to the
B200 implementation.  Same four public names as the reference module (pyviz/apap_utils.py:10,23,40,75)."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from cvx_proj_b200.apap_utils import final_size, get_mesh, get_vertice, uniform_blend  # noqa: E402,F401

__all__ = ["get_mesh", "get_vertice", "final_size", "uniform_blend"]
