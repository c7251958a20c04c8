#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv`: hottest SASS instructions by executed count and by
stall samples, with the main stall reason.  Usage: tools/ncu_hot.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
tot_samp = sum(int(r[ix["# Samples"]]) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total warp instructions", tot_inst, "samples", tot_samp, "sass lines", len(body))
agg = {}
for r in body:
    for s_ in stalls:
        agg[s_] = agg.get(s_, 0) + int(r[ix[s_]] or 0)
print("stall totals:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
ops = {}
for r in body:
    toks = r[ix["Source"]].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
print("opcode mix:", [(k, round(100.0 * v / tot_inst, 1)) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:18]])
print("---- hottest by samples")
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    st = max(stalls, key=lambda s_: int(r[ix[s_]] or 0))
    print("%6s %5.1f%% exec=%9s  %-28s %s" % (r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]]) / max(tot_samp, 1),
                                           r[ix["Instructions Executed"]], st, r[ix["Source"]].strip()[:90]))
