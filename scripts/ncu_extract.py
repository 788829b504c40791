#!/usr/bin/env python3
"""Read a committed ``ncu -i <rep> --page raw --csv`` export and pull out what bounds the trace kernel.

    python scripts/ncu_extract.py profiles/trace_r2_ncu_raw.csv [--kernel rsk_trace_kernel] [--json]

The export has one header row (metric names), one unit row and one row per profiled launch.  ``summary()`` is what
bench.py puts into ``roofline.issue`` / ``roofline.l1`` / ``roofline.l2`` / ``roofline.traffic`` -- every number in
the bench line's roofline block beyond the live CUDA-event timing comes from the file named in ``roofline.source``."""
from __future__ import annotations

import csv
import json
import sys
from pathlib import Path

_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
          "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}


def load(path, kernel: str = "rsk_trace_kernel"):
    """-> list of {metric: (value, unit)} for the launches whose kernel name contains ``kernel``."""
    rows = list(csv.reader(open(path, newline="")))
    if len(rows) < 3:
        raise ValueError(f"{path}: not a raw-page export")
    names, units = rows[0], rows[1]
    k_col = names.index("Kernel Name")
    out = []
    for r in rows[2:]:
        if kernel in r[k_col]:
            out.append({n: (v, u) for n, u, v in zip(names, units, r)})
    return out


def _num(launch, name, default=None):
    if name not in launch:
        return default
    v, u = launch[name]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return default
    return x * _SCALE.get(u, 1.0)


def summary(path, kernel: str = "rsk_trace_kernel", rays: int | None = None) -> dict:
    launches = load(path, kernel)
    if not launches:
        raise ValueError(f"{path}: no launch of {kernel}")
    k = launches[0]
    g = lambda n, d=None: _num(k, n, d)     # noqa: E731
    issue = g("smsp__issue_active.avg.pct_of_peak_sustained_active")
    lanes = g("smsp__thread_inst_executed_per_inst_executed.ratio")
    warp_inst = g("smsp__inst_executed.sum")
    l1_sectors = g("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
    l1_req = g("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")
    l1_wave = g("l1tex__data_pipe_lsu_wavefronts_mem_lg_cmd_read.sum", g("l1tex__data_pipe_lsu_wavefronts_mem_lg.sum"))
    l2_sectors = g("lts__t_sectors_srcunit_tex_op_read.sum", g("lts__t_sectors_op_read.sum"))
    dram = (g("dram__bytes_read.sum", 0.0) or 0.0) + (g("dram__bytes_write.sum", 0.0) or 0.0)
    s = {
        "source": str(path),
        "kernel": k["Kernel Name"][0],
        "kernel_ms": g("gpu__time_duration.sum"),
        "registers_per_thread": g("launch__registers_per_thread"),
        "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue": {"issue_active_pct": issue, "lanes_per_inst": lanes,
                  "useful_lane_issue_frac": None if issue is None or lanes is None else issue / 100.0 * lanes / 32.0,
                  "warp_instructions": warp_inst,
                  "pipe_alu_pct": g("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                  "pipe_fma_pct": g("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                  "pipe_lsu_pct": g("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                  "pipe_xu_pct": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")},
        "l1": {"bytes": None if l1_sectors is None else 32.0 * l1_sectors,
               "pct_of_peak": g("l1tex__throughput.avg.pct_of_peak_sustained_active"),
               "hit_pct": g("l1tex__t_sector_hit_rate.pct"),
               "global_load_requests": l1_req, "global_load_sectors": l1_sectors, "global_load_wavefronts": l1_wave},
        "l2": {"bytes": None if l2_sectors is None else 32.0 * l2_sectors,
               "pct_of_peak": g("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
               "hit_pct": g("lts__t_sector_hit_rate.pct")},
        "dram": {"bytes": dram, "pct_of_peak": g("dram__throughput.avg.pct_of_peak_sustained_elapsed")},
        "sm_throughput_pct": g("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    }
    if rays:
        s["per_ray"] = {"warp_instructions": None if warp_inst is None else warp_inst / rays,
                        "l1_bytes": None if l1_sectors is None else 32.0 * l1_sectors / rays,
                        "dram_bytes": dram / rays}
    return s


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    kern = "rsk_trace_kernel"
    if "--kernel" in sys.argv:
        kern = sys.argv[sys.argv.index("--kernel") + 1]
    rays = int(sys.argv[sys.argv.index("--rays") + 1]) if "--rays" in sys.argv else 239026176
    print(json.dumps(summary(Path(sys.argv[1]), kern, rays), indent=1))
