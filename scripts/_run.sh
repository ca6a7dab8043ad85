set -x
timeout 900 python -m pytest tests/test_gpu_codes.py -m gpu -x -q -k "streaming or block_length or large_block" > gpurun_out/r2g_tests.log 2>&1; tail -15 gpurun_out/r2g_tests.log
for rule in f64 f64fused fast; do timeout 300 python scripts/stream_probe.py 100000 51080 9472 0.10 20 $rule; done > gpurun_out/r2g_probe.log 2>&1
timeout 300 python scripts/stream_probe.py 10240 5231 9472 0.10 20 f64 3 >> gpurun_out/r2g_probe.log 2>&1
cat gpurun_out/r2g_probe.log
