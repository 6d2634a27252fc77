"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches_c2.md
    python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_fused_kernel_c2.md
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def launches(src, dst):
    lines = [l for l in open(src) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row["Kernel Name"].split("(")[0], []).append(float(row["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | mean us | share of captured GPU time |\n|---|---|---|---|\n")
        for k, v in agg.items():
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, len(v), sum(v) / len(v) / 1e3, 100 * sum(v) / total))
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            f.write("### %s\n\n| metric | value | unit |\n|---|---|---|\n" % d.get("Kernel Name", ("", "?"))[1].split("(")[0])
            for k in KEYS:
                if k in d:
                    f.write("| %s | %s | %s |\n" % (k, d[k][1], d[k][0]))
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
