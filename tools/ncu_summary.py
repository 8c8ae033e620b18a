"""Turns an .ncu-rep (ncu --set full) into a compact per-kernel table, and an ncu launch list
(gpu__time_duration.sum CSV) into per-step shares.  Run in the build container:

    python tools/ncu_summary.py kernels gpurun_out/prof.ncu-rep > profiles/xxx.md
    python tools/ncu_summary.py launches gpurun_out/launches.csv [marker-kernel] > profiles/yyy.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit smem (blocks)"),
    ("launch__occupancy_limit_registers", "occupancy limit regs (blocks)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
]


def kernels(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{rep.split('/')[-1]}` (per launch, cold caches, clocks not locked)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        print(f"## {name}\n")
        print("| metric | value | unit |\n|---|---|---|")
        for k, label in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                print(f"| {label} (`{k}`) | {d[k]} | {units[hdr.index(k)]} |")
        print()


def launches(path, marker="fdr_project"):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[1:] if r[idi].isdigit()]
    marks = [i for i, (k, _) in enumerate(data) if marker in k]
    print(f"# ncu launch list `{path.split('/')[-1]}`: {len(data)} launches, {len(marks)} steps "
          f"(a step starts at `{marker}`)\n")
    if len(marks) < 2:
        return
    a, b = marks[-2], marks[-1]
    step = data[a:b]
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for k, v in step:
        k = re.sub(r"<.*", "", k)[:80]
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    print(f"One step = {len(step)} kernel launches, {tot / 1e3:.1f} us of GPU time summed over launches "
          "(ncu serialises launches and flushes caches: compare SHARES, not absolutes).\n")
    print("| GPU time (us) | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {v / 1e3:.1f} | {100 * v / tot:.1f} % | {n} | `{k}` |")


if __name__ == "__main__":
    {"kernels": kernels, "launches": launches}[sys.argv[1]](*sys.argv[2:])
