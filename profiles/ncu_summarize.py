"""Summarise an ncu capture exported with `ncu -i X.ncu-rep --page raw --csv > raw.csv` and
`--page source --csv > src.csv`: headline metrics, instruction counts per SASS region, stall reasons."""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
for r in rows[2:3]:
    print(r[hdr.index('Kernel Name')][:80])
    for w in want:
        if w in hdr:
            print(f"  {w:85s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
rows = list(csv.reader(open(src)))
b = rows[1:]
hdr, data = b[0], b[1:]
isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
iw, iwi = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
groups = []
for i, r in enumerate(data):
    if not r[iex]:
        continue
    e, s, w, wi = int(r[iex]), int(r[ismp]), int(r[iw] or 0), int(r[iwi] or 0)
    if groups and groups[-1]['e'] == e and groups[-1]['end'] == i - 1:
        g = groups[-1]; g['end'] = i; g['n'] += 1; g['s'] += s; g['w'] += w; g['wi'] += wi
    else:
        groups.append({'start': i, 'end': i, 'e': e, 'n': 1, 's': s, 'first': r[isrc], 'w': w, 'wi': wi})
tot = sum(g['e'] * g['n'] for g in groups); tots = sum(g['s'] for g in groups)
print("total warp instructions", tot, "stall samples", tots)
for g in groups:
    share = g['e'] * g['n'] / tot * 100
    if share > 0.7 or g['s'] > tots * 0.012:
        print(f"{g['start']:5d}-{g['end']:5d} n={g['n']:3d} exec={g['e']:9d} instr={g['e']*g['n']/1e6:6.2f}M {share:5.1f}% samples={g['s']:4d} "
              f"wavefronts={g['w']} ideal={g['wi']}  {g['first'].strip()[:48]}")
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_')]
t = {}
for r in data:
    for i in stall:
        if r[i]:
            t[hdr[i]] = t.get(hdr[i], 0) + int(r[i])
ss = sum(t.values())
print({k: round(100 * v / ss, 1) for k, v in sorted(t.items(), key=lambda x: -x[1])[:12]})
