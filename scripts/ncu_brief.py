#!/usr/bin/env python
"""ncu_brief.py rep [top_n] -- headline counters, stall mix and the hottest source lines of one capture."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread ', 'launch__occupancy_limit', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ',
        'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ', 'launch__waves']
for h, u, v in zip(hdr, units, vals):
    if any((h + ' ').startswith(w) for w in want):
        print(f"{h} [{u}] {v}")
st = {h.split('stalled_')[1]: int(v.replace(',', '')) for h, v in zip(hdr, vals) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h}
tot = sum(st.values()) or 1
print("stalls:", ", ".join(f"{k} {100*v/tot:.0f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; data = rows[hi + 1:]
ci = h.index("Warp Stall Sampling (All Samples)"); ii = h.index("Instructions Executed"); si = h.index("Source")
tot_s = sum(int(r[ci]) for r in data if len(r) > ci and r[ci].isdigit()) or 1
print(f"SASS instructions: {len(data)}; samples {tot_s}")
# correlate with CUDA source through a second dump
src2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; per = collections.Counter(); ins = collections.Counter()
for r in csv.reader(src2.splitlines()):
    if len(r) >= 2 and r[0].isdigit() and not r[0].startswith("0x") and len(r) < 6:
        pass
for line in src2.splitlines():
    pass
top = sorted(data, key=lambda r: -(int(r[ci]) if len(r) > ci and r[ci].isdigit() else 0))[:topn]
for r in top:
    print(f"{100*int(r[ci])/tot_s:5.1f}%  exec {r[ii]:>10}  {r[si].strip()[:90]}")
