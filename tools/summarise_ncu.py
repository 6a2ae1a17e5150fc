"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  python tools/summarise_ncu.py launches gpurun_out/launches_cfg5.csv profiles/r1_launches_bench_cfg5.txt "<header>"
  python tools/summarise_ncu.py full gpurun_out/pcg_cfg5.ncu-rep profiles/r1_ncu_full_cfg5_kernels.txt "<header>"
"""
import csv
import collections
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def launches(src, dst, header):
    rows = [r for r in csv.reader(open(src, errors="replace")) if r and r[0] and r[0][0].isdigit() or (r and r[0] == "ID")]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1e-6)
        name = r[ik].split("(")[0][:70]
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + ms)
    total = sum(t for _, t in tot.values())
    with open(dst, "w") as f:
        f.write(f"# {header}\n# per-launch times are cold-cache and serialised (ncu replay); compare SHARES, not absolute times\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total_ms':>12s} {'avg_ms':>9s} {'share':>7s}\n")
        for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name:70s} {n:8d} {t:12.3f} {t / n:9.4f} {100 * t / total:6.1f}%\n")
    print(open(dst).read())


def full(src, dst, header):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    with open(dst, "w") as f:
        f.write(f"# {header}\n")
        f.write("kernel: " + " | ".join(r[ik].split("(")[0][:40] for r in data) + "\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"{m} [{units[i]}]: " + " | ".join(r[i] for r in data) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
