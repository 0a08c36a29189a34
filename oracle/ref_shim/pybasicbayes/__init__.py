"""ORACLE shim: minimal stand-in for the six pybasicbayes symbols the reference imports
(pyglm/regression.py:35-36, pyglm/networks.py:8-9, pyglm/models.py:2).  pybasicbayes is a
third-party PyPI package (unpinned in setup.py:13, upstream 0.2.x); its behaviour is restated
from its published semantics -- see SURVEY.md Appendix B.2/B.3.  PARITY UNPINNED for the
random draws (no reference test pins them)."""
