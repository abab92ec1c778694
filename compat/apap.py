"""Drop-in module name of the reference's ``pyviz/apap.py``: ``from apap import APAP`` / ``import apap`` give the
B200 implementation of ``class APAP`` (pyviz/apap.py:21-217) with the reference's constructor and methods, and --
like the reference module, which star-imports its helpers (pyviz/apap.py:15) -- the ``apap_utils`` names."""
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)

from apap_utils import *  # noqa: E402,F401,F403  (pyviz/apap.py:15)
from cvx_proj_b200.apap import APAP  # noqa: E402,F401
from cvx_proj_b200.driver import mat_layout, save2mat  # noqa: E402,F401  (the driver's post-processing, pyviz/apap.py:250-265)
