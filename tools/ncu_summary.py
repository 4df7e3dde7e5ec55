#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + per-line stall samples): python tools/ncu_summary.py rep [topN]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed',
        'l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'sm__cycles_active.avg']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for w in WANT:
        if w in d:
            print('  %-78s %s %s' % (w, d[w], units[hdr.index(w)]))
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page concatenates kernels: split on "Kernel Name" rows
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}
        blocks.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
for b in blocks:
    h = b["hdr"]
    ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot = {s: 0 for s in stalls}
    lines = []
    for r in b["rows"]:
        if len(r) < len(h):
            continue
        try:
            ns = int(r[ix['# Samples']])
        except ValueError:
            continue
        lines.append((ns, r))
        for s in stalls:
            try:
                tot[s] += int(r[ix[s]])
            except ValueError:
                pass
    print("==", b["name"][:100])
    print("   stalls:", {k: v for k, v in sorted(tot.items(), key=lambda x: -x[1]) if v > 0})
    lines.sort(key=lambda x: -x[0])
    for ns, r in lines[:topn]:
        st = {s.replace("stall_", ""): int(r[ix[s]]) for s in stalls if r[ix[s]] not in ('', '0')}
        st = dict(sorted(st.items(), key=lambda x: -x[1])[:3])
        print("   %6d  %-70s %s" % (ns, r[ix['Source']][:70], st))
