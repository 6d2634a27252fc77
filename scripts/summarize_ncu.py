"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r1_launches_c2.md
    python scripts/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r1_fused_kernel_c2.md
    python scripts/summarize_ncu.py multi gpurun_out/launches_c3_gemm.csv profiles/r1_launches_c3_gemm.md
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
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
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


def multi(src, dst):
    """Launch list captured with several --metrics: one row per kernel, one column per metric (means)."""
    lines = [l for l in open(src) if l.startswith('"')]
    agg = collections.OrderedDict()
    units = {}
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        v = row["Metric Value"].replace(",", "")
        if v == "":
            continue
        agg.setdefault(k, collections.OrderedDict()).setdefault(row["Metric Name"], []).append(float(v))
        units[row["Metric Name"]] = row["Metric Unit"]
    metrics = list(units)
    with open(dst, "w") as f:
        f.write("| kernel | launches | " + " | ".join("%s [%s]" % (m, units[m]) for m in metrics) + " |\n")
        f.write("|---|---|" + "---|" * len(metrics) + "\n")
        for k, d in agg.items():
            n = max(len(v) for v in d.values())
            f.write("| `%s` | %d | " % (k, n) + " | ".join("%.4g" % (sum(d[m]) / len(d[m])) if m in d else "" for m in metrics) + " |\n")
    print(open(dst).read())


if __name__ == "__main__":
    {"launches": launches, "full": full, "multi": multi}[sys.argv[1]](sys.argv[2], sys.argv[3])
