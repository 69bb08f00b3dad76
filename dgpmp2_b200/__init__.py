"""dgpmp2_b200 -- B200-native implementation of dGPMP2's inner Gauss-Newton loop.

Layout:
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/dgpmp2_b200.h)
  _lib.py, ops.py  ctypes binding / tensor-level entry points
  gpmp2/ robot_models/ utils/ env/ datasets/
                   host-side mirror of the reference's diff_gpmp2 API for this path
The top-level ``diff_gpmp2`` package of this repository re-exports these modules under the
reference's import paths so the reference's example scripts run unchanged.
"""
__version__ = '0.1.0'
