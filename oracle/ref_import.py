"""Import the UNMODIFIED reference (pvjosue/CWFA at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (fixture generation) and by
oracle validation here in the build container.  /root/reference does not exist on the GPU
box, so nothing that runs there imports this file.

The reference needs a few non-arithmetic third-party modules that are not installed
(matplotlib, tifffile, multipagetiff, lion_pytorch) and one numpy-1 path
(numpy.lib.arraysetops).  None of them carries arithmetic of the hot path; they are stubbed.
Recipe: SURVEY.md section 8(c).
"""
import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("CWFA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "networks.py"))


def import_reference():
    """Returns (networks, CWFA, Ff, Fm, INN_utils) modules of the reference."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    sys.dont_write_bytecode = True
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm",
              "tifffile", "multipagetiff", "lion_pytorch"):
        if m not in sys.modules:
            sys.modules[m] = MagicMock()
    if "numpy.lib.arraysetops" not in sys.modules:
        import numpy as np
        shim = types.ModuleType("numpy.lib.arraysetops")
        shim.isin = np.isin
        sys.modules["numpy.lib.arraysetops"] = shim
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import FrEIA.framework as Ff
    import FrEIA.modules as Fm
    import INN_utils
    import networks
    try:
        import CWFA
    except Exception:  # CWFA.py pulls in plotting/tensorboard; not needed for arithmetic
        CWFA = None
    return networks, CWFA, Ff, Fm, INN_utils
