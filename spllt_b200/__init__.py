"""spllt_b200: B200-native numerical phase of SpLLT behind the reference's C ABI.

The package is a thin ctypes mirror of include/spllt_iface.h + include/spllt_b200.h; all
compute lives in libspllt_b200.so (hand-written sm_100a CUDA + C++ host analysis).
"""
from .api import (SpLLT, Options, Inform, lib, load_library, ORDER_METIS, ORDER_NATURAL, ORDER_USER,
                  chkerr)  # noqa: F401
from . import matrices  # noqa: F401
