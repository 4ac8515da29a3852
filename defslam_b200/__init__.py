"""defslam_b200 -- B200-native implementation of DefSLAM's deformable hot path.

The product is the CUDA library ``libdefslam_b200.so`` (sources in ``csrc/``) behind
the C ABI of ``include/defslam_b200.h``.  This Python package is only the thin
test/bench harness around that ABI (ctypes bindings + synthetic workloads).
"""
from . import _capi  # noqa: F401

__all__ = ["_capi"]
__version__ = "0.1.0"
