"""Importable alias of the ``mini-sglang_b200`` package (whose directory name is not an identifier).

``import minisgl_b200`` before constructing ``minisgl.llm.LLM`` / parsing server args registers the
``"b200"`` attention backend with the reference.
"""

import importlib as _importlib
import sys as _sys
from pathlib import Path as _Path

_root = str(_Path(__file__).resolve().parent.parent)
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _importlib.import_module("mini-sglang_b200")
globals().update({name: getattr(_pkg, name) for name in _pkg.__all__})
__all__ = list(_pkg.__all__)
PACKAGE = _pkg
