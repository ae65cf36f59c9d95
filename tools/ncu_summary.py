#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + hot-loop instruction mix (reads `ncu -i ... --page raw/source --csv`)."""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
kfilter = sys.argv[2] if len(sys.argv) > 2 else None
def page(p):
    cmd = ["ncu", "-i", rep, "--page", p, "--csv"]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))
rows = page("raw")
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
kn = hdr.index('Kernel Name')
for r in data:
    print("==", r[kn][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w); print("  %-68s %-10s %s" % (w, units[i], r[i]))
    st = [(h.replace('smsp__pcsamp_warps_issue_stalled_', ''), int(r[hdr.index(h)])) for h in hdr
          if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
    tot = sum(v for _, v in st) or 1
    print("  stalls: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(st, key=lambda x: -x[1]) if v * 50 > tot))
rows = page("source")
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
if hdr:
    iS, iE, iSamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = sum(int(r[iE]) for r in data)
    cnt = collections.Counter()
    for r in data: cnt[int(r[iE])] += 1
    print("total warp-instr", tot)
    for e, c in sorted(cnt.items(), key=lambda x: -x[0] * x[1])[:6]:
        print("  exec %9d x %4d instr = %5.1f%%" % (e, c, 100 * e * c / tot))
    mx = max(cnt, key=lambda e: e * cnt[e])
    op = collections.Counter(); samp = collections.Counter()
    for r in data:
        e = int(r[iE])
        if abs(e - mx) > mx * 0.02: continue
        t = r[iS].split(); o = t[1] if t[0].startswith('@') else t[0]
        op[o] += 1; samp[o] += int(r[iSamp])
    print("hot loop (exec %d): %d instr" % (mx, sum(op.values())))
    print("  " + ", ".join("%s %d(%d)" % (o, c, samp[o]) for o, c in op.most_common(30)))
