#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report of the fused kernels per source region.

    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_stalls.py src.csv > profiles/rNN_ncu_stalls.txt
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
sections, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        fp, fn, hdr, j, data = rows[i][1], rows[i + 1][1], rows[i + 2], i + 3, []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            data.append(rows[j]); j += 1
        sections.append((fp, fn, hdr, data)); i = j
    else:
        i += 1
src = open(os.path.join(ROOT, "neighbour_feature_pooling_b200", "csrc", "nfp_stream_impl.cuh")).read().split("\n")


def find(pat):
    for k, line in enumerate(src):
        if pat in line:
            return k + 1
    return None


marks = [("ptx helpers (ldx/stx, mbarrier waits, bulk ops)", find("---- PTX helpers"), find("constexpr int kMaxStages")),
         ("kernel entry, barrier init", find("stream_kernel(const StreamArgs a"), find("// ================================ producer warp")),
         ("producer warp (TMA issue, empty-slot waits)", find("// ================================ producer warp"), find("// ================================ consumer warps")),
         ("consumer prologue (pad zeroing, lane setup)", find("// ================================ consumer warps"), find("// ---- backward, before pass A")),
         ("backward: gy-only stencil part", find("// ---- backward, before pass A"), find("// ---- pass A: per-pixel")),
         ("pass A (FFMA2 dots)", find("// ---- pass A: per-pixel"), find("NFP_STAMP(1);")),
         ("table reduction (wtab -> tfull, inv)", find("NFP_STAMP(1);"), find("// ---- forward value")),
         ("forward value / y store", find("// ---- forward value"), find("// ---- backward: stencil coefficients")),
         ("backward coefficient closure", find("// ---- backward: stencil coefficients"), find("// ---- pass B: gx = stencil")),
         ("pass B (stencil FMAs, staging, TMA bulk store)", find("// ---- pass B: gx = stencil"), find("// ---- host side"))]
print("ncu --set full --import-source on, B200: warp-state samples and executed instructions per source region of")
print("csrc/nfp_stream_impl.cuh (code of inlined helpers is attributed to the helpers' own lines).\n")
for fp, fn, hdr, data in sections:
    if not fp.endswith("nfp_stream_impl.cuh"):
        continue
    idx = {h: c for c, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    per, samp, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    for r in data:
        try:
            ln = int(r[0]); per[ln] += int(r[idx["Instructions Executed"]] or 0); samp[ln] += int(r[idx["# Samples"]] or 0)
        except Exception:
            continue
        for s_ in stalls:
            try:
                st[ln][s_] += int(r[idx[s_]] or 0)
            except Exception:
                pass
    tot, ts = sum(per.values()), sum(samp.values())
    print(fn)
    print(f"  warp instructions executed {tot}, samples {ts}")
    allst = collections.Counter()
    for c in st.values():
        allst.update(c)
    print("  stall reasons, whole kernel: " + ", ".join(f"{n[6:]} {100 * v / max(1, sum(allst.values())):.0f}%" for n, v in allst.most_common(8)))
    for nm, a, b in marks:
        if a is None or b is None:
            continue
        n = sum(v for kk, v in per.items() if a <= kk < b); s_ = sum(v for kk, v in samp.items() if a <= kk < b)
        if n == 0 and s_ == 0:
            continue
        c = collections.Counter()
        for kk, cc in st.items():
            if a <= kk < b:
                c.update(cc)
        top = ", ".join(f"{x[6:]} {v}" for x, v in c.most_common(4))
        print(f"  {nm:50s} instr {100 * n / tot:5.1f}%  samples {100 * s_ / max(ts, 1):5.1f}%   top stalls: {top}")
    print()
