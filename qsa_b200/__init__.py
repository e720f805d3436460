"""Importable alias of the package directory ``quickstart-streaming-agents_b200/``.

The layout contract names the package directory with a hyphen, which Python cannot import directly; this
shim puts that directory on the package search path, so ``import qsa_b200.engine`` loads
``quickstart-streaming-agents_b200/engine.py``.  No code lives here.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "quickstart-streaming-agents_b200")
if not _os.path.isdir(_PKG_DIR):  # pragma: no cover
    raise ImportError(f"package directory missing: {_PKG_DIR}")
__path__.append(_PKG_DIR)
PACKAGE_DIR = _PKG_DIR
__version__ = "0.1.0"
