"""Extracts the metrics the docs quote from `ncu --set full` captures (gpurun_out/*.ncu-rep, scratch) into small CSV
summaries under profiles/ (tracked), and the kernel shares from an ncu launch list.

    python profiles/extract_ncu.py summary gpurun_out/r2_db_scan_kernel.ncu-rep profiles/r2_db_scan_ncu_full_summary.csv
    python profiles/extract_ncu.py launches gpurun_out/r2_db_launches.csv profiles/r2_db_launches_summary.csv
"""
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit", "launch__shared_mem_per_block", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled")


def summary(rep, out):
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(raw.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        w.writerow(["kernel", "", v[h.index("Kernel Name")]])
        for i, name in enumerate(h):
            if any(name.startswith(k) for k in KEEP) and "not_issued" not in name and "pcsamp" not in name:
                w.writerow([name, units[i], v[i]])


def launches(src, out):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    agg = {}
    for r in rows:
        k = r[4].split("(")[0].replace("void unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1])
    tot = sum(a[1] for a in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ns", "share"])
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, a[0], int(a[1]), "%.4f" % (a[1] / tot)])


if __name__ == "__main__":
    {"summary": summary, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
