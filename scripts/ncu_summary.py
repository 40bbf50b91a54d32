#!/usr/bin/env python
"""Summarise `ncu --set full` captures of the codec kernels into a small JSON for profiles/.
usage: ncu_summary.py enc.ncu-rep dec.ncu-rep out.json"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]


def one(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out[k] = {"value": float(vals[i].replace(",", "")), "unit": units[i]}
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, data = rows[1], rows[2:]
    cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    t = collections.Counter()
    for r in data:
        for i in cols:
            try:
                t[hdr[i]] += int(r[i])
            except ValueError:
                pass
    s = sum(t.values()) or 1
    out["stall_breakdown_pct"] = {k: round(100 * v / s, 1) for k, v in t.most_common(8)}
    return out


if __name__ == "__main__":
    json.dump({"encode_kernel": one(sys.argv[1]), "decode_kernel": one(sys.argv[2])}, open(sys.argv[3], "w"), indent=1)
