"""B200-native (sm_100a) implementation of the VarNet + spatial-alignment hot path of
woxuankai/SpatialAlignmentNetwork.  The sub-modules mirror the reference's flat files
(``signal_utils``, ``varnet``, ``unet``, ``cross``, ``ssimloss``, ``lnccloss``, ``miloss``,
``model``) and run on the hand-written CUDA kernels of ``libsan_b200.so`` (C ABI in
``include/san_b200.h``).  There is no CPU or PyTorch-library fallback for the arithmetic."""
__version__ = "0.1.0"
