"""Summarise an .ncu-rep (raw page + SASS page) into the numbers quoted in DESIGN.md / profiles/."""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__cycles_elapsed.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_xu.sum', 'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct',
        'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct', 'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('-----')
    for k in keys:
        if k in d:
            print(f"{k:84s} {d[k]} {units[hdr.index(k)]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(sass.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == 'Address':
        cur['hdr'] = r
    elif cur is not None and cur['hdr'] and len(r) >= len(cur['hdr']) - 2:
        cur['data'].append(r)
for b in blocks:
    h = b['hdr']
    iS, iI, iSm = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    c, cs = Counter(), Counter()
    tot = ts = 0
    for r in b['data']:
        try:
            i, sm = int(r[iI]), int(r[iSm])
        except ValueError:
            continue
        parts = r[iS].split()
        op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
        c[op] += i; cs[op] += sm; tot += i; ts += sm
    print('=====', b['name'][:100], 'warp-instr', tot, 'samples', ts)
    for op, v in c.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
        print(f"  {op:10s} {v / tot * 100:6.2f}% instr {cs[op] / max(ts,1) * 100:6.2f}% samples")
