"""Import alias: `import marl_dmfb_b200` -> the package in ./marl-dmfb_b200/ (hyphenated name)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("marl-dmfb_b200")
sys.modules[__name__] = _pkg
