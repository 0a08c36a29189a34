"""ORACLE (test infrastructure).  Import this BEFORE importing the reference package.

Makes /root/reference's own, unmodified files importable on numpy >= 1.24 and provides
stand-ins for the two third-party packages the reference needs but which are absent here
(setup.py:12-13: pybasicbayes, pypolyagamma; neither vendored nor pinned).  Only used in the
build container to generate tests/golden/ fixtures (oracle/gen_golden.py) and to validate the
numpy restatement in oracle/pyglm_oracle.py; /root/reference does not exist on the GPU box.
"""
import os
import sys
import numpy as np

# numpy aliases removed in 1.24 and still used by the reference
# (pyglm/utils/basis.py:84 np.int, pyglm/regression.py:519 np.float).  np.bool exists again
# on numpy >= 2.0 and must NOT be aliased to the Python bool.
if not hasattr(np, "int"):
    np.int = int
if not hasattr(np, "float"):
    np.float = float

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
REFERENCE_ROOT = os.environ.get("PYGLM_REFERENCE_ROOT", "/root/reference")
if os.path.isdir(REFERENCE_ROOT) and REFERENCE_ROOT not in sys.path:
    sys.path.insert(1, REFERENCE_ROOT)
