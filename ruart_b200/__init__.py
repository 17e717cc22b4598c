"""ruart_b200 — B200-native (sm_100a) implementation of RUArt's per-question inference path.

Package layout
  csrc/            hand-written CUDA kernels + the C ABI (include/ruart_b200.h)
  _lib.py, ops.py  ctypes binding / tensor wrappers (no fallbacks)
  Models/          drop-in mirrors of the reference's Models/Layers.py, Models/SDNet.py,
                   Models/Bert/Bert.py (same classes, constructors and state_dict names)
  Utils/phoc.py    drop-in mirror of the reference's Utils/phoc.py
"""
__version__ = "0.1.0"
