"""Summarise an ncu report (--set full, --import-source on) of one kernel: headline counters and the SASS hot spots.
    python tools/ncu_summary.py gpurun_out/prof_voigt.ncu-rep [region_boundaries...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "launch__registers_per_thread ",
        "launch__occupancy_limit", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum ", "smsp__average_warps_issue_stalled", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_elapsed.max ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum "]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, val = raw[0], raw[1], raw[2]
    print("kernel:", val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, val):
        if any((h + " ").startswith(k) or h.startswith(k) for k in KEYS):
            print("  %-90s %-12s %s" % (h, u, v))
    src = page(rep, "source")
    h = src[1]
    data = src[2:]
    ia, ins, smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ti = sum(int(r[ins]) for r in data)
    ts = sum(int(r[smp]) for r in data)
    print("total warp-instructions %d, samples %d" % (ti, ts))
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
    for k, r in enumerate(data):
        if int(r[ins]) > thr * ti or int(r[smp]) > thr * ts:
            print("%5d %-64s %13s %6.2f%% %8s %6.2f%%" % (k, r[ia][:64], r[ins], 100 * int(r[ins]) / ti, r[smp],
                                                        100 * int(r[smp]) / ts))


if __name__ == "__main__":
    main()
