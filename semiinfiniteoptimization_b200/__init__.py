"""B200-native collision oracle for semi-infinite collision-free optimisation.

Drop-in for the collision-oracle hot path of krishauser/SemiInfiniteOptimization: the reference's
module names are kept (`geometryopt`, `sip`, `constraint`, `objective`); the geometry queries
behind them run as sm_100a CUDA kernels reached through the C-ABI in include/sio_abi.h.
"""
__version__ = "0.1.0"
