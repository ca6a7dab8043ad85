"""qkd_ldpc_b200 -- B200-native (sm_100a) sum-product LDPC reconciliation behind the QKD_LDPC API.

Layout:
  csrc/     hand-written CUDA kernels + the C-ABI (libqkdldpc_b200.so, declared in include/qkd_ldpc_b200.h)
  host/     C++ mirror of the reference's public functions and its config.json-driven simulation, on top of the C-ABI
  capi.py   ctypes binding of the C-ABI (tests, bench.py)
  codes.py  parity-check matrix files (alist / dense / the compact .npz under data/codes)
  sweep.py  frame-batch scheduler for QBER sweeps, sharded over ranks
"""
from .capi import (Code, Context, DecodeParams, QlbError, load_library, make_params, pack_bits, unpack_bits,  # noqa: F401
                   PRECISION_F32, PRECISION_F64, RES_KEYS_MATCH, RES_SYNDROMES_MATCH)

__version__ = "0.1.0"
