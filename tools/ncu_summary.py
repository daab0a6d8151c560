#!/usr/bin/env python3
"""Dev helper: print the key metrics + hottest SASS lines of an .ncu-rep (needs ncu on PATH)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, d = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_active.avg']
print(d[hdr.index('Kernel Name')][:90])
for k in keys:
    if k in hdr:
        i = hdr.index(k); print(f"  {k:75s} {units[i]:16s} {d[i]}")
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(d[i] or 0) > 0.05:
        print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {d[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; n = len(hdr)
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
blk = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) == n: blk.append(r)
tot = sum(int(r[iex]) for r in blk); totS = sum(int(r[isamp]) for r in blk)
print("  total warp-instr %d samples %d" % (tot, totS))
seg = []; cur = None
for r in blk:
    e, s = int(r[iex]), int(r[isamp])
    if cur and abs(cur['e'] - e) <= 0.02 * max(e, 1):
        cur['n'] += 1; cur['tot'] += e; cur['s'] += s; cur['last'] = r[ia][-5:]
    else:
        cur = {'first': r[ia][-5:], 'last': r[ia][-5:], 'e': e, 'n': 1, 'tot': e, 's': s, 'src': r[isrc][:40]}; seg.append(cur)
for s in seg:
    if s['tot'] / tot > 0.004 or s['s'] / totS > 0.01:
        print(f"  {s['first']}-{s['last']} n={s['n']:3d} exec/instr={s['e']:>10d} instr-share={s['tot']/tot*100:5.1f}% samples={s['s']/totS*100:5.1f}%  {s['src']}")
top = sorted(blk, key=lambda r: -int(r[isamp]))[:12]
for r in top:
    print("   hot", r[ia][-5:], f"{int(r[isamp])/totS*100:5.1f}%", r[isrc][:70])
