"""libjxl_b200 -- B200-native JPEG XL VarDCT decode transform pipeline (host-side mirror).

The product is the C-ABI shared library built from libjxl_b200/csrc (include/jxl_b200.h);
this package is the thin Python host layer used by tests and bench.py.
"""
from . import abi  # noqa: F401

__all__ = ["abi"]
