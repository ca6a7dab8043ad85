"""Print the SASS of a kernel between two addresses. usage: sass_range.py <binary> <name-substring> <lo-hex> <hi-hex>"""
import re, subprocess, sys
binary, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", binary], capture_output=True, text=True).stdout
for blk in re.split(r"\n\s*Function : ", out)[1:]:
    name = blk.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    if pat not in dem:
        continue
    for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", blk):
        a = int(m.group(1), 16)
        if lo <= a <= hi:
            print(f"{a:#07x}  {m.group(2).strip()}")
    break
