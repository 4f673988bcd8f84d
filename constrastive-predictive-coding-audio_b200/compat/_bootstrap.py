"""Makes ``cpc_b200`` importable from the shim modules in this directory (they live next to the package root)."""
import os
import sys

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(1, _PKG)
