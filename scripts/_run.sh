for v in "" v1b8 v1b4 v1b8c2 v1b8c4; do
  echo "== variant ${v:-base}"
  if [ -n "$v" ]; then export QLB_LIBRARY=$PWD/qkd_ldpc_b200/lib/variants/lib$v.so; else unset QLB_LIBRARY; fi
  timeout 300 python scripts/stream_probe.py 100000 51080 9472 0.10 20 f64 2>&1 | tail -1
  timeout 300 python scripts/stream_probe.py 100000 51080 9472 0.10 20 f64fused 2>&1 | tail -1
done > gpurun_out/r2y_variants.log 2>&1
cat gpurun_out/r2y_variants.log
export QLB_LIBRARY=$PWD/qkd_ldpc_b200/lib/variants/libv1b8.so
python -m pytest tests/test_gpu_codes.py -m gpu -x -q -k "streaming_fp64 or block_length or large_block" 2>&1 | tail -3
