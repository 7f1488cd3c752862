"""Importable alias of the package directory ``tfc-gan_b200/`` (a hyphen is not a valid module name).

``import tfc_gan_b200 as tfc`` gives the package itself.
"""

import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("tfc-gan_b200")
sys.modules[__name__] = _pkg
