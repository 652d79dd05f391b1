#!/usr/bin/env python
"""Condense .ncu-rep captures into the per-kernel lines kept under profiles/.

usage: ncu_summary.py <out.md> <title> <rep> [<rep> ...]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_rd_sectors"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
    ("smsp__issue_active.avg.per_cycle_active", "ipc/smsp"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dsmem"),
]


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def stalls(hdr, r, top=4):
    """largest warp-stall reasons, in warps stalled per issued instruction"""
    it = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
            if name not in ("selected", "not_selected"):
                it.append((name, float(r[i] or 0)))
    it.sort(key=lambda x: -x[1])
    return ", ".join(f"{k}={v:.2f}" for k, v in it[:top])


def main():
    out, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = [f"# {title}", "",
             "Per-launch figures from `ncu --set full --clock-control none` (cold caches, serialised launches; "
             "durations under the profiler are NOT bench numbers).", ""]
    for rep in reps:
        hdr, units, rows = rows_of(rep)
        lines += [f"## {rep.split('/')[-1]}", "", "| kernel | " + " | ".join(n for _, n in KEYS) + " | top stalls |",
                  "|---|" + "---|" * (len(KEYS) + 1)]
        for r in rows:
            name = r[hdr.index("Kernel Name")]
            vals = []
            for k, _ in KEYS:
                if k in hdr:
                    v, u = r[hdr.index(k)], units[hdr.index(k)]
                    try:
                        v = f"{float(v):.4g}"
                    except ValueError:
                        pass
                    vals.append(f"{v} {u}".strip())
                else:
                    vals.append("-")
            lines.append(f"| `{name[:70]}` | " + " | ".join(vals) + f" | {stalls(hdr, r)} |")
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
