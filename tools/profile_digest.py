#!/usr/bin/env python
"""Digest profiling artefacts into small tracked files under profiles/.

  tools/profile_digest.py sass  LIB.so OUT.txt          per-kernel SASS opcode census (UTMALDG/UTMASTG/UBLKCP/SYNCS ...)
                                                         plus the copy-engine / mbarrier lines of the hot kernels
  tools/profile_digest.py ncu   REPORT.ncu-rep OUT.json  key metrics of every launch in an `ncu --set full` report
  tools/profile_digest.py hot   REPORT.ncu-rep OUT.txt   hottest SASS lines by stall samples (source page)
  tools/profile_digest.py traffic LAUNCHES.csv OUT.json  per-kernel averages of an ncu launch list (time, DRAM bytes, share);
                                                         bench.py reads OUT.json for `roofline.traffic`
"""
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMACMDFLUSH", "LDGSTS", "DFMA", "DMUL", "DADD", "LDS", "STS", "REDUX",
        "BAR", "POPC", "LOP3"]
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "smsp__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg",
]


def sass(lib, out):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    lines = ["# SASS census of %s (cuobjdump -sass; sm_100a only)" % lib, ""]
    excerpts = []
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        ops = {}
        for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)[^;]*;", f):
            op = m.group(1).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
            if op in ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTMACMDFLUSH") and ("k_tile_col" in name or "k_expect_lean" in name or "k_apply_lean" in name):
                excerpts.append("%-28s %s" % (name[:28], m.group(0).strip()))
        lines.append("%-90s total %5d  %s" % (name[:90], sum(ops.values()), {k: ops[k] for k in KEYS if k in ops}))
    lines += ["", "# copy-engine (TMA) and mbarrier instructions in the hot kernels", ""] + excerpts
    open(out, "w").write("\n".join(lines) + "\n")


def ncu_rows(rep, page):
    txt = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(txt)))


def ncu(rep, out):
    rows = ncu_rows(rep, "raw")
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:60]}
        for k in METRICS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = [float(r[i].replace(",", "")), units[i]]
                except ValueError:
                    d[k] = [r[i], units[i]]
        res.append(d)
    json.dump({"report": rep, "launches": res}, open(out, "w"), indent=1)


def hot(rep, out, top=40):
    """Hottest SASS lines by stall samples, with the dominant stall reason, for every distinct launch of the report
    (the source page repeats its header per launch; ncu lists a launch twice when two sections carry the source view)."""
    rows = ncu_rows(rep, "source")
    heads = [i for i, r in enumerate(rows) if "Instructions Executed" in r]
    lines, seen = [], set()
    for li, h in enumerate(heads):
        end = heads[li + 1] - 1 if li + 1 < len(heads) else len(rows)
        hdr = rows[h]
        ix = {c: i for i, c in enumerate(hdr)}
        body = [r for r in rows[h + 1:end] if len(r) == len(hdr)]
        ti = sum(int(r[ix["Instructions Executed"]]) for r in body)
        ts = sum(int(r[ix["# Samples"]]) for r in body)
        key = (ti, ts, len(body))
        if key in seen or not body:
            continue
        seen.add(key)
        stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
        agg = {s_: sum(int(r[ix[s_]] or 0) for r in body) for s_ in stalls}
        title = rows[h - 1][1][:120] if h > 0 and len(rows[h - 1]) > 1 else ""
        lines += ["# launch %d %s" % (len(seen) - 1, title), "# warp instructions %d, stall samples %d, SASS lines %d" % (ti, ts, len(body)),
                  "# stall samples by reason: %s" % [(k[6:], v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]]]
        ops = {}
        for r in body:
            toks = r[ix["Source"]].split()
            op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")).split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
        lines.append("# opcode mix (%% of executed warp instructions): %s" %
                     [(k, round(100.0 * v / max(ti, 1), 1)) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:16]])
        lines += ["", "samples   share  executed  main stall        instruction"]
        for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
            st = max(stalls, key=lambda s_: int(r[ix[s_]] or 0))
            lines.append("%7s %6.1f%% %9s  %-16s  %s" % (r[ix["# Samples"]], 100.0 * int(r[ix["# Samples"]]) / max(ts, 1),
                                                         r[ix["Instructions Executed"]], st[6:], r[ix["Source"]].strip()[:100]))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")


def traffic(launches_csv, out):
    import gzip
    op = gzip.open if launches_csv.endswith(".gz") else open
    with op(launches_csv, "rt") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = {}
    for r in rows:
        d = per.setdefault(r["ID"], {"name": re.sub(r"^void ", "", r["Kernel Name"]).split("<")[0].split("(")[0]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * ({"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
                                                                         if r["Metric Name"].startswith("gpu__time") else 1.0)
    agg = {}
    for d in per.values():
        agg.setdefault(d["name"], []).append(d)
    total = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
    res = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                     "`python bench.py --steps 1 --warmup 1 --no-cpu --no-pool` (%s), averages per launch" % launches_csv}
    for name, ds in agg.items():
        t = [d.get("gpu__time_duration.sum", 0.0) for d in ds]
        res[name] = {"launches": len(ds), "avg_us": sum(t) / len(t), "min_us": min(t), "max_us": max(t),
                     "dram_read_bytes": sum(d.get("dram__bytes_read.sum", 0.0) for d in ds) / len(ds),
                     "dram_write_bytes": sum(d.get("dram__bytes_write.sum", 0.0) for d in ds) / len(ds),
                     "share": sum(t) / max(total, 1e-12)}
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"sass": sass, "ncu": ncu, "hot": hot, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
