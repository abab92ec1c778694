"""Recipe for ``oracle/_ref/``: the reference's own APAP scripts, unmodified, next to the oracle so that
``bench.py --impl reference`` can time the REAL reference (``APAP.local_homography`` / ``local_warp`` of
``pyviz/apap.py``) on the GPU box's host cores, where ``/root/reference`` does not exist.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under ``cvx_proj_b200/`` may import it.

    python oracle/make_ref.py          (run by ``__graft_entry__.build()`` when /root/reference is present)

copies ``pyviz/{apap,apap_utils,utils,baseline_stitch_test}.py`` (MIT, see /root/reference/LICENSE; apap.py imports the
other three at module level) byte for byte into ``oracle/_ref/pyviz/``.  ``oracle/_ref/`` is git-ignored: the files
never enter this repository's history; they travel to the GPU box with the gpurun snapshot like a built ``.so``.
Without them the reference arm falls back to the oracle port and says so (``cpu_baseline.kind = "port"``).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pyviz"
DST = os.path.join(HERE, "_ref", "pyviz")
FILES = ("apap.py", "apap_utils.py", "utils.py", "baseline_stitch_test.py")


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"oracle/make_ref: {SRC} not present, nothing to do")
        return False
    os.makedirs(DST, exist_ok=True)
    for name in FILES:
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    lic = "/root/reference/LICENSE"
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(DST, "LICENSE"))
    with open(os.path.join(DST, "SHA256SUMS"), "w") as f:
        for name in FILES:
            f.write(hashlib.sha256(open(os.path.join(DST, name), "rb").read()).hexdigest() + "  " + name + "\n")
    if verbose:
        print(f"oracle/make_ref: {len(FILES)} reference files -> {DST}")
    return True


def load():
    """Import the vendored reference (``np.int = int`` is the only shim: apap_utils.py:59 uses the alias numpy
    removed).  Returns the reference's ``apap`` module (its ``apap_utils`` as attribute ``apap_utils_module``) or None
    when ``oracle/_ref/pyviz`` is absent.  ``sys.modules`` is left as it was."""
    if not all(os.path.exists(os.path.join(DST, n)) for n in FILES):
        return None
    import importlib

    import numpy as np
    if not hasattr(np, "int"):
        np.int = int  # noqa
    names = ("apap", "apap_utils", "utils", "baseline_stitch_test")
    saved = {k: sys.modules.pop(k, None) for k in names}
    sys.path.insert(0, DST)
    try:
        mod = importlib.import_module("apap")
        mod.apap_utils_module = importlib.import_module("apap_utils")
        return mod
    finally:
        sys.path.remove(DST)
        for k in names:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]


if __name__ == "__main__":
    make()
