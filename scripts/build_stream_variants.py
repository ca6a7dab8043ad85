"""Kernel-tuning helper: rebuilds ONE translation unit (default: the streaming one) with different -D knobs and links each
against the other objects of the library into qkd_ldpc_b200/lib/variants/lib<tag>.so; select one at run time with
QLB_LIBRARY=<path>. usage: build_stream_variants.py [--tu qlb_tu_resident_f32.cu] tag=FLAGS ...   e.g.  c3="-DQLB_SPLIT_CHECK_MINB=3" """
import subprocess, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from qkd_ldpc_b200 import build as b

TU = "qlb_tu_stream.cu"
if len(sys.argv) > 2 and sys.argv[1] == "--tu":
    TU = sys.argv[2]
    del sys.argv[1:3]
b.build_library()
out = b.LIB_DIR / "variants"
out.mkdir(exist_ok=True)
others = [b.LIB_DIR / "obj" / (Path(t).stem + ".o") for t in b.TRANSLATION_UNITS if t != TU]

def one(spec):
    tag, flags = spec.split("=", 1)
    obj = out / f"{tag}.o"
    cmd = [b.NVCC, *b.NVCC_FLAGS, *flags.split(), "-c", "-o", str(obj), str(b.CSRC / TU), "-Xptxas", "-v"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    (out / f"{tag}.log").write_text(p.stdout + p.stderr)
    if p.returncode:
        return f"{tag}: compile failed\n{p.stderr[-2000:]}"
    p = subprocess.run([b.NVCC, *b.NVCC_FLAGS, "-shared", "-o", str(out / f"lib{tag}.so"), str(obj), *map(str, others), "-ldl"], capture_output=True, text=True)
    obj.unlink(missing_ok=True)
    return f"{tag}: {'ok' if p.returncode == 0 else p.stderr[-2000:]}"

with ThreadPoolExecutor(max_workers=4) as pool:
    for r in pool.map(one, sys.argv[1:]):
        print(r, flush=True)
