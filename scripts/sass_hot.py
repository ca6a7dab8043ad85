"""Hot regions of a kernel from an ncu source page (sass): contiguous runs of instructions that executed at least `thr` times the
maximum, with their opcode mix and stall samples.  usage: sass_hot.py <sass.csv> [thr=0.2] [dump_region_index]"""
import csv, sys
from collections import Counter
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
dump = int(sys.argv[3]) if len(sys.argv) > 3 else -1
hdr = rows[1]
iS, iI, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    try:
        data.append((r[iS].strip(), int(r[iI]), int(r[iSm]), r))
    except (ValueError, IndexError):
        pass
mx = max(d[1] for d in data)
tot = sum(d[1] for d in data)
regions, cur = [], None
for k, d in enumerate(data):
    if d[1] >= thr * mx:
        if cur is None or k - cur[1] > 8:
            cur = [k, k]
            regions.append(cur)
        cur[1] = k
print(f"total warp instructions {tot}, max per instruction {mx}, {len(data)} SASS lines")
for ri, (a, b) in enumerate(regions):
    seg = data[a:b + 1]
    n = sum(d[1] for d in seg)
    sm = sum(d[2] for d in seg)
    ops = Counter()
    st = Counter()
    for d in seg:
        parts = d[0].split()
        op = parts[1] if parts and parts[0].startswith('@') else (parts[0] if parts else '?')
        ops[op.split('.')[0]] += d[1]
        for i, h in stall_cols:
            try:
                st[h] += int(d[3][i])
            except ValueError:
                pass
    print(f"region {ri}: lines {a}-{b} ({b - a + 1} instr), {n} executed ({100 * n / tot:.1f} %), samples {sm}, per-instr exec ~{seg[len(seg)//2][1]}")
    print("   ops: " + ", ".join(f"{o} {100 * c / n:.1f}%" for o, c in ops.most_common(14)))
    print("   stalls: " + ", ".join(f"{h[6:]} {100 * c / max(1, sum(st.values())):.1f}%" for h, c in st.most_common(8)))
    if ri == dump:
        for d in seg:
            print(f"{d[1]:>12} {d[2]:>7}  {d[0]}")
