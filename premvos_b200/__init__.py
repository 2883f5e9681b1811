"""premvos_b200: B200-native (sm_100a) implementation of the PReMVOS per-frame dense compute hot path.

The compute lives in premvos_b200/lib/libpremvos_b200.so (hand-written CUDA behind the C ABI in
include/premvos_b200.h).  This package is the host-side mirror of the reference's Python call
surface.  Importing it never loads the CPU oracle and never falls back to PyTorch ops."""
from . import _lib  # noqa: F401
from .pwc import (Correlation, PWCDCNet, calculate_flow, correlation_forward, pwc_dc_net, readFlowFile,  # noqa: F401
                  writeFlowFile)

__version__ = "0.1"
