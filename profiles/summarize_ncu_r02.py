#!/usr/bin/env python
"""Round-2 ncu artefacts (gpurun_out/, produced by tests/tools/evidence_run.sh) -> tracked summaries:
  r02_ncu_launches_resident{0,1}.csv : gpu__time_duration per launch of `bench.py --steps 3 --warmup 3 --no-sweep --no-e2e`
  r02_ncu_pull.ncu-rep               : `ncu --set full` of sw_pull_kernel, one launch per batch shape (`sw_probe pull 1`)
Writes profiles/r02_ncu_summary.md and updates profiles/traffic.json."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
here = os.path.dirname(os.path.abspath(__file__))
out = ["# ncu summary (round 2)\n"]


def launch_table(name, title):
    path = os.path.join(src, name)
    if not os.path.exists(path):
        return
    agg = collections.OrderedDict()
    lines = [l for l in open(path) if not l.startswith("==")]
    for row in csv.DictReader(lines):
        try:
            k, v = row["Kernel Name"].split("(")[0], float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    ours = {k: v for k, v in agg.items() if k.startswith("sw_")}
    tot = sum(v[1] for v in ours.values()) or 1.0
    out.append(f"## Launch list, {title}\n")
    out.append("`ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 python bench.py --steps 3 --warmup 3 "
               "--no-sweep --no-e2e --no-cpu-baseline` (serialised, cold-cache launches: compare SHARES, not absolutes; "
               "under ncu the resident kernels cannot overlap, so each lives until its idle linger or lifetime ends)\n")
    out.append("| kernel | launches | total us | avg us | share of our kernel time |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | {t / tot * 100:.1f} % |")
    other = {k: v for k, v in agg.items() if not k.startswith("sw_")}
    if other:
        n, t = sum(v[0] for v in other.values()), sum(v[1] for v in other.values())
        out.append(f"| (torch kernels of the harness: buffer fills, payload check) | {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | - |")
    out.append("")


launch_table("r02_ncu_launches_resident1.csv", "resident engine (default)")
launch_table("r02_ncu_launches_resident0.csv", "discrete-kernel engine (`STARWAY_RESIDENT=0`)")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum"]
SHAPES = [(1, 1 << 20), (8, 1 << 20), (16, 1 << 20), (64, 1 << 20), (64, 4 << 20), (16, 64 << 20), (1, 1 << 30), (52, 16384), (52, 65536)]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1, "msecond": 1e6}
rep = os.path.join(src, "r02_ncu_pull.ncu-rep")
traffic = {}
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out.append("## `ncu --set full --clock-control none --import-source on -k regex:sw_pull_kernel -c 9 sw_probe pull - 1` (SW_PROBE_PULL_LINGER_US=1)\n")
    out.append("One launch per batch shape of tests/gpu_probe/probe.cu (the batch is published first, the kernel then "
               "claims it, copies it, writes its completion records and leaves on the EXIT batch). Under ncu the kernel "
               "duration includes start-up, the descriptor prefetch and the completion records, not only the copy.\n")
    out.append("| batch | payload | grid x block | regs | dyn smem | duration us | DRAM read | DRAM write | DRAM bytes / payload byte | DRAM % of peak | payload GB/s over the whole launch |\n|---|---|---|---|---|---|---|---|---|---|---|")
    for i, r in enumerate(data[: len(SHAPES)]):
        def val(m):
            if m not in col:
                return float("nan")
            v = float(r[col[m]].replace(",", "") or "nan")
            return v * UNIT.get(units[col[m]], 1)
        nmsg, ln = SHAPES[i]
        pay = nmsg * ln
        dur = val("gpu__time_duration.sum")
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        out.append(f"| {nmsg} x {ln >> 10} KiB | {pay / 1e6:.1f} MB | {r[col['launch__grid_size']]} x {r[col['launch__block_size']]} | "
                   f"{r[col['launch__registers_per_thread']]} | {r[col['launch__shared_mem_per_block_dynamic']]} {units[col['launch__shared_mem_per_block_dynamic']]} | "
                   f"{dur / 1e3:.1f} | {rd / 1e6:.1f} MB | {wr / 1e6:.1f} MB | {(rd + wr) / pay:.3f} | "
                   f"{r[col['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]} | {pay / dur:.0f} |")
        traffic[f"{nmsg}x{ln}"] = {"payload_bytes": pay, "dram_read": rd, "dram_write": wr, "duration_ns": dur}
    out.append("")
    out.append("| batch | instructions executed | warps active % of peak | L2 sector hit rate % | occupancy limit (shared memory) | waves per SM |\n|---|---|---|---|---|---|")
    for i, r in enumerate(data[: len(SHAPES)]):
        nmsg, ln = SHAPES[i]
        g = lambda m: r[col[m]] if m in col else "-"  # noqa: E731
        out.append(f"| {nmsg} x {ln >> 10} KiB | {g('smsp__inst_executed.sum')} | {g('sm__warps_active.avg.pct_of_peak_sustained_active')} | "
                   f"{g('lts__t_sector_hit_rate.pct')} | {g('launch__occupancy_limit_shared_mem')} CTA/SM | {g('launch__waves_per_multiprocessor')} |")
    out.append("")
    out.append("(One elected thread per CTA drives the TMA pipeline -- `UBLKCP` global->shared->global, 8 x 24 KiB stages -- and a "
               "second warp prefetches the next batch's descriptors: ~3 % of the warp slots are active by design; the bytes "
               "move through the TMA unit, not through registers.  76 % of DRAM peak *over the whole launch* at 1 GiB "
               "includes ~20 us of start-up, descriptor fetch and completion records.)")
    out.append("")
    out.append("Algorithmic traffic is 2 bytes of DRAM per payload byte (one read, one write); a ratio near 2.0 means no "
               "wasted re-reads. Sources smaller than the 126 MB L2 that were just written by the probe's memset can "
               "read below 1 byte per byte from DRAM.\n")
    if "--source" in sys.argv:
        pass

open(os.path.join(here, "r02_ncu_summary.md"), "w").write("\n".join(out) + "\n")
tj = os.path.join(here, "traffic.json")
old = json.load(open(tj)) if os.path.exists(tj) else {}
if traffic:
    old["r02_sw_pull_kernel"] = traffic
    json.dump(old, open(tj, "w"), indent=1)
print("\n".join(out))
