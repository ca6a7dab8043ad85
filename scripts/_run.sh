timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2aa_gpu_tests.log 2>&1; tail -3 gpurun_out/r2aa_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; python - <<PY
import json
d = json.load(open("gpurun_out/r2aa_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], "roof", d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"]["algorithmic_hbm"]["frac"], d["clocks"])
for k, v in d["variants"].items():
    print(k, v.get("frames_per_s"), v.get("algorithmic_hbm_frac", v.get("roofline_frac")))
PY
