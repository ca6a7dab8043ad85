timeout 900 python -m pytest tests/test_gpu_codes.py -m gpu -x -q -k "streaming" > gpurun_out/r2l_tests.log 2>&1; tail -3 gpurun_out/r2l_tests.log
for rule in f64 f64fused fast; do timeout 300 python scripts/stream_probe.py 100000 51080 9472 0.10 20 $rule | tail -1; done > gpurun_out/r2l_probe.log 2>&1
timeout 600 python scripts/stream_probe.py 1000000 510800 4096 0.10 12 f64 | tail -1 >> gpurun_out/r2l_probe.log 2>&1
cat gpurun_out/r2l_probe.log
