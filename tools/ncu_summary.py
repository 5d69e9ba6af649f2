"""Turns ncu CSV exports into the markdown summaries kept under profiles/.

  launches : python tools/ncu_summary.py launches <launch_list.csv> "<command>" > profiles/rNN_launches.md
             (csv from `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> <command>`)
  raw      : python tools/ncu_summary.py raw <raw.csv> "<command>" [kernel-substring] > profiles/rNN_ncu_<kernel>.md
             (csv from `ncu -i report.ncu-rep --page raw --csv > raw.csv` of an `ncu --set full` capture)

Both run on the CPU box: the .ncu-rep stays on the GPU box (too large for gpurun_out), only the CSV travels."""
import csv
import io
import sys
from collections import OrderedDict

RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
]


def read_csv(path):
    lines = [l for l in open(path, newline="") if l.startswith('"')]
    return list(csv.reader(io.StringIO("".join(lines))))


def short(name: str) -> str:
    name = name.replace("void ", "").replace("jb::", "")
    return name.split("(")[0][:56]


def launches(path, command):
    rows = read_csv(path)
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = OrderedDict()
    total = 0.0
    n = 0
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        k = short(r[ki])
        c, t = per.get(k, (0, 0.0))
        per[k] = (c + 1, t + us)
        total += us
        n += 1
    print(f"# ncu launch list of `{command}`\n")
    print(f"`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> {command}` - per-launch times are "
          "cold-cache and serialised: read the shares, not the absolutes. Under CUDA injection the library keeps one launch "
          "per round (the persistent tail kernel cannot be replayed by a profiler).\n")
    print(f"{n} launches, {total:.0f} us of device time.\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {c} | {t:.1f} | {100 * t / total:.1f} % |")


def raw(path, command, want=None):
    rows = read_csv(path)
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full: `{command}`\n")
    print("Values from `ncu -i ... --page raw --csv` (per launch; ncu serialises and replays kernels, so compare shares, not absolutes).\n")
    for r in rows[2:]:
        if want and want not in r[ki]:
            continue
        print(f"## {r[ki]}  grid={r[hdr.index('Grid Size')]} block={r[hdr.index('Block Size')]}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for m in RAW_METRICS:
            cols = [i for i, h in enumerate(hdr) if h == m or h.endswith("." + m)]
            if cols:
                print(f"| {m} | {r[cols[0]]} | {units[cols[0]]} |")
        stalls = []
        for i, h in enumerate(hdr):
            if "smsp__pcsamp_warps_issue_stalled_" in h and not h.endswith("_not_issued") and r[i] not in ("", "0"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h.split("stalled_")[1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls)
        if tot:
            top = sorted(stalls, reverse=True)[:6]
            print("\nTop stall reasons (pc sampling): " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in top))
        print()


if __name__ == "__main__":
    if len(sys.argv) < 4 or sys.argv[1] not in ("launches", "raw"):
        sys.exit(__doc__)
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        raw(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
