"""Import alias: `import smg_b200` == the package in `symbolic-music-generation_b200/` (hyphenated names cannot be
written in an `import` statement)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('symbolic-music-generation_b200')
sys.modules[__name__] = _pkg
