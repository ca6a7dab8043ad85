#!/bin/bash
# Builds lib variants/libbounds.so with -DQLB_BOUNDS_CHECK (every kernel tests the indices it is about to use and traps on the
# first one out of range) and runs the cross-kernel agreement script and the streaming / parity tests against it.
# compute-sanitizer is closed on the B200 pool (profiles/r02_compute_sanitizer_closed.log); this is its stand-in.
#   here (CPU box):  bash scripts/bounds_check_build.sh build
#   on the GPU box:  bash scripts/bounds_check_build.sh run > gpurun_out/bounds.log
set -e
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  for tu in qlb_api qlb_tu_resident_f32 qlb_tu_resident_f64 qlb_tu_stream; do
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ccbin /usr/bin/g++ -Xcompiler -fPIC -DQLB_BOUNDS_CHECK -DQLB_STREAM_LEAN \
      -c -o qkd_ldpc_b200/lib/variants/bounds_$tu.o qkd_ldpc_b200/csrc/$tu.cu &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o qkd_ldpc_b200/lib/variants/libbounds.so qkd_ldpc_b200/lib/variants/bounds_*.o -ldl
  rm -f qkd_ldpc_b200/lib/variants/bounds_*.o
  ls -la qkd_ldpc_b200/lib/variants/libbounds.so
else
  export QLB_LIBRARY=$PWD/qkd_ldpc_b200/lib/variants/libbounds.so
  python scripts/sanitize_case.py
  python -m pytest tests/test_gpu_codes.py tests/test_gpu_parity.py -m gpu -q -x -k "(streaming and not other_bit_weights) or golden or waterfall_fp64 or edge_cases or no_clamp or multi_rate or storage_tiers or block_length"
fi
