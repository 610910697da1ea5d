#!/usr/bin/env python3
"""Markdown table of the bench lines in a directory (gpurun_out/<tag>/ or profiles/): for DESIGN.md section 5.

    python tools/bench_table.py gpurun_out/r2_final [prefix]
"""
import glob
import json
import os
import sys

ORDER = ["cfg5", "cfg2", "cfg4a", "cfg3", "cfg2p", "cfg6ch", "cfg4b", "cfg4c"]
WHAT = {"cfg5": "cfg5: 65,536 × stereo × 1 s (default)", "cfg2": "cfg2: 1,024 × stereo × 10 s", "cfg4a": "cfg4a: 4,096 × 8 ch × 2 s",
        "cfg3": "cfg3: 16,384 × mono × 320 frames, 50 ticks per span launch", "cfg2p": "cfg2p: cfg2 × 5 s + float planes (8 B/sample)",
        "cfg6ch": "cfg6ch: 4,096 × 6 ch × 1 s", "cfg4b": "cfg4b (EXTENSION, unpinned): 8→2 downmix, 10 channels metered",
        "cfg4c": "cfg4c (EXTENSION, unpinned): 8→2 downmix, outputs metered"}


def main():
    d = sys.argv[1]
    prefix = sys.argv[2] if len(sys.argv) > 2 else "bench_"
    print("| workload | kernel | ms per step | Msamples/s | of 6,545.6 GB/s | sustained ≥ 2 s | in place, no overlap | end to end Msamples/s (of link) |")
    print("|---|---|---|---|---|---|---|---|")
    for w in ORDER:
        f = os.path.join(d, f"{prefix}{w}.json")
        if not os.path.exists(f):
            continue
        x = json.loads(open(f).read().strip().splitlines()[-1])
        r = x["roofline"]
        e = x.get("e2e") or {}
        sus = r.get("sustained", {}).get("frac")
        ip = r.get("in_place_no_overlap", {}).get("frac")
        e2e = f"{e['value'] / 1e3:.1f}e3 ({e['frac_of_link']:.2f})" if e.get("frac_of_link") else "—"
        print(f"| {WHAT[w]} | `{r['kernel']}` | {x['ms_per_step']:.3f} | {x['value'] / 1e6:.3f}e6 | **{r['frac']:.3f}** | "
              f"{sus:.3f} | {'—' if ip is None else f'{ip:.3f}'} | {e2e} |" if sus is not None else
              f"| {WHAT[w]} | `{r['kernel']}` | {x['ms_per_step']:.3f} | {x['value'] / 1e6:.3f}e6 | **{r['frac']:.3f}** | — | — | {e2e} |")


if __name__ == "__main__":
    main()
