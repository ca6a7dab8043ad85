"""profiles/traffic.json from ncu reports (--set full, one launch each) of the bench's own launch shape.
usage: make_traffic.py key=report.ncu-rep:frame_iterations_of_that_launch:bytes_per_edge_iteration:edges[:capture text] ...
Per key it records: DRAM bytes of the launch (dram__bytes_read.sum + dram__bytes_write.sum), warp instructions and FP64-pipe warp
instructions (SASS opcodes D*) per frame-iteration, kernel time, and the summary file written beside it under profiles/."""
import csv, json, subprocess, sys
from collections import Counter
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
out_path = ROOT / "profiles" / "traffic.json"
data = json.loads(out_path.read_text()) if out_path.exists() else {}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
for spec in sys.argv[1:]:
    key, rest = spec.split("=", 1)
    parts = rest.split(":")
    rep, frame_it, bpe, edges = parts[0], int(parts[1]), int(parts[2]), int(parts[3])
    text = parts[4] if len(parts) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    def val(name):
        v, u = d[name]
        return float(v.replace(",", "")) * UNIT.get(u, 1)
    dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    winst = val("smsp__inst_executed.sum")
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(sass.splitlines()))
    h = next(r for r in srows if r and r[0] == "Address")
    iS, iI = h.index("Source"), h.index("Instructions Executed")
    ops = Counter()
    for r in srows:
        try:
            n = int(r[iI])
        except (ValueError, IndexError):
            continue
        p = r[iS].split()
        if not p:
            continue
        op = (p[1] if p[0].startswith("@") and len(p) > 1 else p[0]).split(".")[0]
        ops[op] += n
    fp64 = sum(n for op, n in ops.items() if op in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
    mufu = ops.get("MUFU", 0)
    entry = {
        "dram_bytes_per_launch": int(dram),
        "algorithmic_bytes_of_that_launch": frame_it * edges * bpe,
        "capture": f"{Path(rep).name}: {d['Kernel Name'][0][:90]}, {frame_it} frame-iterations in the launch; {text}".strip("; "),
        "kernel_ms": val("gpu__time_duration.sum") * (1e-6 if d["gpu__time_duration.sum"][1] in ("ns", "nsecond") else 1e-3 if d["gpu__time_duration.sum"][1] in ("us", "usecond") else 1.0),
        "warp_instructions_per_frame_iteration": round(winst / frame_it, 1),
        "fp64_warp_instructions_per_frame_iteration": round(fp64 / frame_it, 1),
        "mufu_warp_instructions_per_frame_iteration": round(mufu / frame_it, 1),
        "fp64_lanes_per_sm_per_clock": 58.0,
        "fp64_lanes_source": "scripts/micro/fp64_peak.cu on B200: 16.9 T DFMA/s = 58 lanes per SM and clock sustained (nominal 64); profiles/r02_fp64_peak.log",
        "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
        "fp64_pipe_active_pct": float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
        "registers": int(float(d["launch__registers_per_thread"][0])), "block": int(float(d["launch__block_size"][0])),
    }
    data[key] = entry
    print(key, json.dumps(entry)[:400])
out_path.write_text(json.dumps(data, indent=1) + "\n")
