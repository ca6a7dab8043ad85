"""Static look at the loops of a kernel in a .so/.cubin: body length and opcode mix (no GPU needed).
usage: sass_loops.py <binary> <mangled-or-substring-of-kernel-name> [min_body]"""
import re, subprocess, sys
from collections import Counter
binary, pat = sys.argv[1], sys.argv[2]
min_body = int(sys.argv[3]) if len(sys.argv) > 3 else 20
names = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", names)
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if pat not in dem and pat not in name:
        continue
    ins = []
    for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", blk):
        ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr_index = {a: i for i, (a, _) in enumerate(ins)}
    print("==", dem[:120], "instructions:", len(ins))
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA\S*\s+(?:.*\s)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    for lo, hi in sorted(set(loops)):
        body = ins[lo:hi + 1]
        if len(body) < min_body:
            continue
        ops = Counter()
        for _, t in body:
            parts = t.split()
            op = parts[1] if parts[0].startswith("@") else parts[0]
            ops[op.split(".")[0]] += 1
        tag = "MUFU x%d" % ops["MUFU"] if ops["MUFU"] else ("VOTE" if ops["VOTE"] else "")
        print(f"  loop {ins[lo][0]:#x}-{ins[hi][0]:#x} len={len(body):4d} {tag:10s} {dict(ops.most_common(12))}")
