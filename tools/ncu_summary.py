#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, on CPU, with `ncu -i`) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_fused_tick_cfg2.txt [workload-key]

Writes the per-launch key metrics, the warp-stall breakdown and the top stalled SASS lines, and
(with a workload key) records dram bytes per launch in profiles/traffic.json for bench.py.
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], Path(sys.argv[2])
    key = sys.argv[3] if len(sys.argv) > 3 else None
    raw = ncu_csv(rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = [f"# summary of {Path(rep).name} (ncu --set full --clock-control none --import-source on)", ""]
    traffic = None
    for li, r in enumerate(raw[2:]):
        name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        lines.append(f"## launch {li}: {name}")
        vals = {}
        for i, h in enumerate(hdr):
            if h in KEYS:
                vals[h] = (r[i], units[i])
        for k in KEYS:
            if k in vals:
                lines.append(f"{k:75s} {vals[k][0]:>16s} {vals[k][1]}")
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(vals["dram__bytes_read.sum"][0]) * scale[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0]) * scale[vals["dram__bytes_write.sum"][1]]
            traffic = rd + wr
            lines.append(f"{'dram bytes read+write per launch':75s} {traffic:16.0f} byte")
        except Exception:
            pass
        stalls = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        lines.append("warp stall samples: " + ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(stalls, reverse=True)[:7]))
        lines.append("")
    src = ncu_csv(rep, "source")
    if len(src) > 2:
        h = src[1]
        ia, isrc, isamp, ilsb = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("stall_long_sb")
        rows = []
        for r in src[2:]:
            try:
                rows.append((int(r[isamp]), int(r[ilsb] or 0), r[ia][-5:], r[isrc]))
            except (ValueError, IndexError):
                pass
        tot = sum(x[0] for x in rows) or 1
        lines.append(f"## top stalled SASS lines (first launch; {tot} samples)")
        for s, l, a, t in sorted(rows, reverse=True)[:12]:
            lines.append(f"{100 * s / tot:5.1f}%  long_sb={l:6d}  ...{a}  {t.strip()[:100]}")
    dst.write_text("\n".join(lines) + "\n")
    if key and traffic:
        tj = dst.parent / "traffic.json"
        d = json.loads(tj.read_text()) if tj.exists() else {}
        d[key] = {"dram_bytes_per_launch": traffic, "from": dst.name}
        tj.write_text(json.dumps(d, indent=1) + "\n")
    print(dst.read_text())


if __name__ == "__main__":
    main()
