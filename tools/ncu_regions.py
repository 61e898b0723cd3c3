"""Stall samples and executed instructions of a voigt_tile_kernel capture, aggregated by phase of the kernel.
The phases are located by marker lines in the CUDA source embedded in the report (--import-source on), so the
mapping follows the source the capture was taken with.
    python tools/ncu_regions.py gpurun_out/prof_voigt_l50.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

MARKERS = [("line state (S, widths, wing)", "struct LineState"),
           ("closed form / reciprocal / Weideman series", "__device__ __forceinline__ double rcp_nr"),
           ("records, searches, TIPS", "struct __align__(16) WideRec"),
           ("kernel set-up (TMA tile, registers)", "__global__ void __launch_bounds__"),
           ("prologue: records and masks", "auto produce = "),
           ("pass 1: visit set-up", "auto consume = "),
           ("pass 1: tile-level far-field visits (32 nodes per tile)", "// tile-level far-field visits:"),
           ("pass 1: far-field visits (32 nodes per slice)", "// far-field visits:"),
           ("pass 1: test-free visits", "unsigned m = s_mfull[buf][blk][warp];"),
           ("pass 1: partial visits", "m = s_mtouch["),
           ("pass 2: wings and cores", "---- pass 2"),
           ("chunk loop and barriers", "if (!WS) {"),
           ("far-field epilogue (coefficients, Clenshaw)", "---- far-field epilogue"),
           ("(after the kernel)", "// exact eval counter")]
RAW = ["gpu__time_duration.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "launch__grid_size", "launch__registers_per_thread"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def report(rep):
    rows, cur, hdr = [], None, None
    for r in csv.reader(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass").splitlines()):
        if not r:
            continue
        if r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
        elif hdr is not None and r[0].isdigit():
            rows.append((cur, int(r[0]), r[1], I(r[i_s]), I(r[i_i])))
    src, cur = {}, None                                   # full listing of voigt.cu as embedded in the report
    for r in csv.reader(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda").splitlines()):
        if r and r[0] in ("File Path", "File Name"):
            cur = r[1].split("/")[-1]
        elif r and cur == "voigt.cu" and r[0].isdigit():
            src[int(r[0])] = r[1]
    starts = []
    for name, marker in MARKERS:
        hit = [ln for ln in sorted(src) if marker in src[ln]]
        if hit:
            starts.append((hit[0], name))
    starts.sort()
    ts, ti = sum(r[3] for r in rows), sum(r[4] for r in rows)
    raw = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    d = dict(zip(raw[0], raw[2]))
    out = ["## %s" % rep.split("/")[-1], "kernel: %s" % d.get("Kernel Name", "?")]
    out += ["%-70s %s" % (k, d[k]) for k in RAW if k in d]
    out.append("%-46s %9s %9s" % ("phase", "%samples", "%instr"))
    for k, (ln0, name) in enumerate(starts):
        ln1 = starts[k + 1][0] if k + 1 < len(starts) else 10 ** 9
        sm = sum(r[3] for r in rows if r[0] == "voigt.cu" and ln0 <= r[1] < ln1)
        ins = sum(r[4] for r in rows if r[0] == "voigt.cu" and ln0 <= r[1] < ln1)
        if sm or ins:
            out.append("%-46s %8.1f%% %8.1f%%" % (name, 100.0 * sm / ts, 100.0 * ins / ti))
    sm = sum(r[3] for r in rows if r[0] != "voigt.cu")
    ins = sum(r[4] for r in rows if r[0] != "voigt.cu")
    out.append("%-46s %8.1f%% %8.1f%%" % ("inlined headers (shuffles, ballots, math)", 100.0 * sm / ts, 100.0 * ins / ti))
    return "\n".join(out)


if __name__ == "__main__":
    print("\n\n".join(report(r) for r in sys.argv[1:]))
