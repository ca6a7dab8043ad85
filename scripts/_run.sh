N=$1
python -m pytest tests/test_host_cpp.py -m gpu -q -x 2>&1 | tail -2
python scripts/run_sim.py 1000000 32fast 1 $N > gpurun_out/r2s_sim_${N}gpu_f32fast.txt 2>&1; head -2 gpurun_out/r2s_sim_${N}gpu_f32fast.txt
python scripts/run_sim.py 1000000 32fast 1 1 > gpurun_out/r2s_sim_1of${N}gpu_f32fast.txt 2>&1; head -2 gpurun_out/r2s_sim_1of${N}gpu_f32fast.txt
python scripts/run_sim.py 1000000 64 1 $N > gpurun_out/r2s_sim_${N}gpu_f64.txt 2>&1; head -2 gpurun_out/r2s_sim_${N}gpu_f64.txt
cmp <(tail -n +3 gpurun_out/r2s_sim_${N}gpu_f32fast.txt) <(tail -n +3 gpurun_out/r2s_sim_1of${N}gpu_f32fast.txt) && echo "CSV identical on $N and 1 GPU (f32fast)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --no-variants --strong-frames-per-point 1000000 > gpurun_out/r2s_bench_${N}gpu.json 2> gpurun_out/r2s_bench_${N}gpu.err; python - <<PY
import json
d = json.load(open("gpurun_out/r2s_bench_${N}gpu.json"))
print("bench N=${N}: value", d["value"], "e2e", d["e2e"]["value"], "strong", d["strong"]["value"], d["strong"]["ms"])
PY
