"""lqcd_b200 -- Python twin of the Julia shim over liblqcd_b200.so (B200-native Dirac-solve path)."""
from . import _lib
from ._lib import LqcdError, NotConverged, WILSON, STAGGERED, OP_D, OP_DDAG, OP_DDAGD  # noqa: F401
from .api import *  # noqa: F401,F403
