#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

  python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.md
  python scripts/ncu_summary.py full gpurun_out/prof_blend.ncu-rep profiles/rNN_blend_full.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = rows[0]
    i_name, i_metric, i_val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    i_unit = hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        v = v / 1e3 if r[i_unit] in ("ns", "nsecond") else (v * 1e3 if r[i_unit] in ("ms", "msecond") else v)  # -> us
        name = r[i_name].split("(")[0][-70:]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: {src}; {sum(a[0] for a in agg.values())} launches, {tot / 1e3:.3f} ms total\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of {src} (--clock-control none)\n\n")
        seen = set()
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0]
            if name in seen:
                continue
            seen.add(name)
            f.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in idx:
                    f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
            f.write("\n")
            _record_traffic(name, r, idx, units, src)
    print(open(dst).read())


_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def _record_traffic(name, r, idx, units, src):
    """profiles/traffic.json: DRAM bytes (read + write) of one launch per kernel, the figure bench.py reports as
    roofline.traffic.  Keyed by the kernel name up to its first template argument (e.g. blend_bwd_gp_kernel<17)."""
    import json
    import os
    import re
    if "dram__bytes_read.sum" not in idx:
        return
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[idx[k]].replace(",", "")) * _UNIT.get(units[idx[k]], 1.0)
    m = re.match(r"(?:void )?(?:d4::)?(\w+)<(?:\((?:int|bool)\))?(\d+)", name)
    key = f"{m.group(1)}<{m.group(2)}" if m else name
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    num = lambda k: float(r[idx[k]].replace(",", "")) if k in idx else None
    d[key] = {"dram_bytes_per_launch": tot, "ms": num("gpu__time_duration.sum"),
              "source": os.path.basename(src), "grid": r[idx["launch__grid_size"]] if "launch__grid_size" in idx else None,
              # fp32-issue roofline of the blend kernels (SURVEY 8d asks for both rooflines)
              "inst_executed": num("smsp__inst_executed.sum"),
              "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
              "shared_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")}
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
