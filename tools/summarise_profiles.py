"""Turn the ncu outputs of tools/gpu_profile.sh (gpurun_out/) into the tracked summaries under profiles/.
    python tools/summarise_profiles.py r01"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def kernel_summary(rep, tag, round_tag):
    raw = page(rep, "raw")
    hdr, units, val = raw[0], raw[1], raw[2]
    d = dict(zip(hdr, val))
    u = dict(zip(hdr, units))
    lines = ["# ncu --set full --clock-control none --import-source on, one launch of %s" % d.get("Kernel Name", tag),
             "# command: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e  (config 2, 1 GPU)", ""]
    for k in RAW_KEYS:
        if k in d:
            lines.append("%-88s %-14s %s" % (k, u[k], d[k]))
    src = page(rep, "source")
    h = src[1]
    data = src[2:]
    ia, ins, smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ti = sum(int(r[ins]) for r in data)
    ts = sum(int(r[smp]) for r in data)
    lines += ["", "# SASS hot spots (>= 1%% of warp instructions or of stall samples); totals: %d instructions, %d samples" % (ti, ts),
              "# %5s %-60s %14s %7s %9s %7s" % ("idx", "instruction", "executed", "%inst", "samples", "%smp")]
    for k, r in enumerate(data):
        if int(r[ins]) > 0.01 * ti or int(r[smp]) > 0.01 * ts:
            lines.append("  %5d %-60s %14s %6.2f%% %9s %6.2f%%" % (k, r[ia][:60], r[ins], 100 * int(r[ins]) / ti, r[smp],
                                                                 100 * int(r[smp]) / ts))
    mn = {}
    for r in data:
        op = r[ia].split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
        op = op.split(".")[0]
        mn[op] = mn.get(op, 0) + int(r[ins])
    lines += ["", "# executed warp instructions by opcode (top 14)"]
    for op, c in sorted(mn.items(), key=lambda kv: -kv[1])[:14]:
        lines.append("  %-12s %14d %6.2f%%" % (op, c, 100 * c / ti))
    open(os.path.join(OUT, "%s_%s_ncu_full.txt" % (round_tag, tag)), "w").write("\n".join(lines) + "\n")
    return d


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    pre = sys.argv[2] if len(sys.argv) > 2 else ""            # file-name prefix inside gpurun_out/ (round 2: "r2_")
    os.makedirs(OUT, exist_ok=True)
    # launch list
    rows = [r for r in csv.reader(open(os.path.join(GP, pre + "launches.csv"))) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = {}
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)",
           "# command: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-banded", "# id, kernel, ns"]
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        ns = float(r[iv].replace(",", ""))
        out.append("%s, %s, %.0f" % (r[0], name, ns))
        tot[name] = tot.get(name, 0.0) + ns
    s = sum(tot.values())
    out.append("# share of device time by kernel")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        out.append("#   %-70s %10.3f ms %6.2f%%" % (k, v / 1e6, 100 * v / s))
    open(os.path.join(OUT, "%s_launches.csv" % tag), "w").write("\n".join(out) + "\n")
    dv = kernel_summary(os.path.join(GP, pre + "prof_voigt.ncu-rep"), "voigt_tile_kernel", tag)
    dt = kernel_summary(os.path.join(GP, pre + "prof_tud.ncu-rep"), "tud_kernel", tag)

    def mb(d, k):
        return float(d[k].replace(",", ""))
    traffic = {"source": "%s ncu --set full, per launch" % tag,
               "voigt_tile_kernel_dram_bytes_per_launch": 1e6 * (mb(dv, "dram__bytes_read.sum") + mb(dv, "dram__bytes_write.sum")),
               "tud_kernel_dram_bytes_per_launch": 1e6 * (mb(dt, "dram__bytes_read.sum") + mb(dt, "dram__bytes_write.sum")),
               "units_note": "ncu reports Mbyte for these launches",
               "voigt_tile_kernel_fp64_pipe_pct": mb(dv, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               "voigt_tile_kernel_issue_active_pct": mb(dv, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "voigt_tile_kernel_ms_under_ncu": mb(dv, "gpu__time_duration.sum"),
               "tud_kernel_fp64_pipe_pct": mb(dt, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               "tud_kernel_issue_active_pct": mb(dt, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "tud_kernel_ms_under_ncu": mb(dt, "gpu__time_duration.sum")}
    json.dump(traffic, open(os.path.join(OUT, "roofline_traffic.json"), "w"), indent=1)
    print(open(os.path.join(OUT, "%s_launches.csv" % tag)).read()[-900:])


if __name__ == "__main__":
    main()
