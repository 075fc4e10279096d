#!/usr/bin/env python
"""Turns the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/:
  launches_sw.csv (gpu__time_duration per launch of our kernels during `bench.py --steps 5 --warmup 3`)
  prof_bulk / prof_match / prof_put .ncu-rep (`ncu --set full`)
Writes r01_ncu_summary.md and traffic.json (DRAM bytes per payload byte of the bulk kernel)."""
import collections
import csv
import json
import os
import subprocess
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
here = os.path.dirname(os.path.abspath(__file__))
out = ["# ncu summary (round 1)\n"]

# ---- launch list
agg = collections.OrderedDict()
path = os.path.join(src, "launches_sw.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    for row in csv.DictReader(lines):
        try:
            k, v = row["Kernel Name"].split("(")[0], float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values()) or 1.0
    out.append("## Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^sw_ ... python bench.py --steps 5 --warmup 3`\n")
    out.append("(cold-cache, serialised launches: compare SHARES, not absolutes)\n")
    out.append("| kernel | launches | total us | avg us | share of our kernel time |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | {t / tot * 100:.1f} % |")
    out.append("")

# ---- full captures
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum"]
traffic = None
for name in ("prof_bulk", "prof_match", "prof_put"):
    rep = os.path.join(src, name + ".ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    out.append(f"## `ncu --set full` — {name}\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        out.append(f"**{d.get('Kernel Name', '?').split('(')[0]}** grid {d.get('launch__grid_size')} x block {d.get('launch__block_size')}")
        for w in WANT:
            if w in d:
                out.append(f"- {w} = {d[w]} {units[hdr.index(w)]}")
        out.append("")
        if name == "prof_bulk":
            def num(key):
                v = float(d[key].replace(",", ""))
                u = units[hdr.index(key)].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            traffic = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
with open(os.path.join(here, "r01_ncu_summary.md"), "w") as f:
    f.write("\n".join(out) + "\n")
print("\n".join(out))
if traffic is not None and len(sys.argv) > 2:
    payload = float(sys.argv[2])  # payload bytes of the captured bulk launch
    with open(os.path.join(here, "traffic.json"), "w") as f:
        json.dump({"world": 1, "kernel": "sw_bulk_tma_kernel", "dram_bytes": traffic, "payload_bytes": payload,
                   "dram_bytes_per_payload_byte": traffic / payload,
                   "source": "ncu --set full, last captured launch (dram__bytes_read.sum + dram__bytes_write.sum)"}, f)
