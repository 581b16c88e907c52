"""monopsr_b200 -- B200-native (sm_100a) implementation of MonoPSR's per-instance hot path.

Host code is Python with torch tensors as device-memory containers only; all compute is
hand-written CUDA in ``monopsr_b200/csrc`` reached through the C ABI declared in
``include/*.h`` (ctypes).  There is NO CPU fallback: importing an op module without the
built library, or calling an op with a non-CUDA tensor, raises.
"""
__version__ = "0.1"
