python -m pytest tests/test_gpu_codes.py -m gpu -x -q -k "streaming or block_length" 2>&1 | tail -3
for lanes in 1 2; do
for rule in f64 f64fused; do timeout 300 python scripts/stream_probe.py 100000 51080 9472 0.10 20 $rule - $lanes | tail -2; done
done > gpurun_out/r2w_probe.log 2>&1
timeout 300 python scripts/stream_probe.py 100000 51080 9472 0.085 100 f64 - 1 | tail -1 >> gpurun_out/r2w_probe.log 2>&1
timeout 300 python scripts/stream_probe.py 100000 51080 9472 0.085 100 f64 - 2 | tail -1 >> gpurun_out/r2w_probe.log 2>&1
timeout 600 python scripts/stream_probe.py 1000000 510800 4096 0.10 12 f64 - 2 | tail -1 >> gpurun_out/r2w_probe.log 2>&1
timeout 300 python scripts/stream_probe.py 100000 51080 18944 0.10 20 fast - 2 | tail -1 >> gpurun_out/r2w_probe.log 2>&1
cat gpurun_out/r2w_probe.log
