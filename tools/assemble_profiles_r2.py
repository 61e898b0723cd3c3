"""Assemble the tracked round-2 summaries under profiles/ from the outputs of tools/gpu_profile_r2b.sh and
tools/gpu_profile_r2c.sh in gpurun_out/.    python tools/assemble_profiles_r2.py"""
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")


def rd(name):
    return open(os.path.join(O, name)).read()


def launches(name):
    rows = list(csv.reader(open(os.path.join(O, name))))
    hdr = [r for r in rows if r and r[0] == "ID"][0]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    out = []
    for r in rows:
        if r and r[0].isdigit():
            n = re.sub(r"<unnamed>::", "", r[ik]); n = re.sub(r"^void ", "", n); n = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", n)
            out.append((int(r[0]), n, r[ig], r[iv]))
    return out


def grab(text, key):
    for ln in text.splitlines():
        if ln.strip().startswith(key):
            return float(ln.split()[-1])


def main():
    open(os.path.join(P, "r02_bench_1gpu.json"), "w").write(rd("r2c_bench_1gpu.json"))
    with open(os.path.join(P, "r02_launches.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  (cold-cache, serialised: compare SHARES)\n"
                "# command: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-banded   (final build of round 2: tools/gpu_profile_r2c.sh)\n"
                "# the tile kernel appears three times per step: K = 8 without / with tile-level far-field blocks and K = 4; every state's\n"
                "# CTAs run in exactly one of them (chosen on the device), the others retire at their first instruction\n"
                "# id, kernel, ns\n")
        for i, n, g, v in launches("r2c_launches.csv"):
            f.write("%d, %s, %s\n" % (i, n, v))
    v, t = rd("r2b_voigt_summary.txt"), rd("r2b_tud_summary.txt")
    open(os.path.join(P, "r02_voigt_tile_kernel_ncu_full.txt"), "w").write(
        "# ncu --set full --clock-control none --import-source on, one launch of voigt_tile_kernel<8, 1, 0, 0, 0> (K = 8, one-step reciprocal, "
        "no tile-level blocks: the variant every state of config 2 below layer ~30 takes)\n# command: python bench.py --steps 2 --warmup 1 "
        "--no-cpu-baseline --no-e2e --no-banded --flags 18432  (config 2, 1 GPU, all 66 layers pinned to this variant; tools/gpu_profile_r2b.sh, final build)\n\n" + v)
    open(os.path.join(P, "r02_tud_kernel_ncu_full.txt"), "w").write(
        "# ncu --set full --clock-control none --import-source on, one launch of tud_kernel\n# command: python bench.py --steps 2 --warmup 1 "
        "--no-cpu-baseline --no-e2e --no-banded  (config 2, 1 GPU; tools/gpu_profile_r2b.sh, final build)\n\n" + t)
    json.dump({"source": "round 2 (final build) ncu --set full, per launch; voigt_tile_kernel = the K = 8 instantiation with all 66 layers of config 2 "
                         "pinned to it (tools/gpu_profile_r2b.sh)",
               "voigt_tile_kernel_dram_bytes_per_launch": (grab(v, "dram__bytes_read.sum") + grab(v, "dram__bytes_write.sum")) * 1e6,
               "tud_kernel_dram_bytes_per_launch": (grab(t, "dram__bytes_read.sum") + grab(t, "dram__bytes_write.sum")) * 1e6,
               "units_note": "ncu reports Mbyte for these launches",
               "voigt_tile_kernel_fp64_pipe_pct": grab(v, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               "voigt_tile_kernel_issue_active_pct": grab(v, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "voigt_tile_kernel_ms_under_ncu": grab(v, "gpu__time_duration.sum"),
               "tud_kernel_fp64_pipe_pct": grab(t, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
               "tud_kernel_issue_active_pct": grab(t, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "tud_kernel_ms_under_ncu": grab(t, "gpu__time_duration.sum")}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
    # regimes: whole-launch section replaced, layer groups kept
    rp = os.path.join(P, "r02_voigt_regimes.txt")
    s = open(rp).read()
    i, j = s.index("## r2b_prof_voigt.ncu-rep"), s.index("# Layer-group sections below")
    new = rd("r2b_regimes_all.txt").strip().replace(
        "## r2b_prof_voigt.ncu-rep", "## r2b_prof_voigt.ncu-rep  (final build of round 2; all 66 layers of config 2 pinned to K = 8 without tile-level blocks, --flags 18432)")
    open(rp, "w").write(s[:i] + new + "\n\n" + s[j:])
    # microbench file: first two sections regenerated, the rest kept
    mp = os.path.join(P, "r02_voigt_microbench.txt")
    s = open(mp).read()
    k = s.index("# tools/xs_breakdown.py")
    head = ("# tools/voigt_microbench.py 5 on the final round-2 build (one B200; ms per rtx_voigt_sum call, tile size pinned by flags)\n"
            "# flags: 0 default (one-step reciprocal, slice- and tile-level far-field blocks), 16384 no tile-level blocks, 128 exact reciprocal,\n"
            "#        4096 no far-field blocks, 2 FP32 asymptote branch, 4 warp-specialised\n" + rd("r2b_voigt_microbench.txt") +
            "\n# tools/tilefar_sweep.py: layers 0-7 of config 2 (210k lines over 800 cm^-1), every line's wing set to OmegaWing; 'automatic' = variant picked on the device\n" +
            rd("r2b_tilefar_sweep.txt") + "\n")
    open(mp, "w").write(head + s[k:])
    # config 1
    ls = [x for x in launches("r2c_cfg1_launches.csv") if 40 <= x[0] < 44]
    open(os.path.join(P, "r02_cfg1_tiles.txt"), "w").write(
        "# tools/cfg1_run.py on the final build of round 2: config 1 (one state, 50 001 points, 20 000 lines), us per rtx_voigt_sum call\n"
        "# over 300 back-to-back C-ABI calls; K = 0 is the automatic choice (512-point tiles, lines split over 7 CTAs per tile), a forced K switches the split off\n" +
        rd("r2c_cfg1_tiles.txt") + "\n# per kernel of the automatic choice (ncu --metrics gpu__time_duration.sum, cold and serialised):\n" +
        "".join("%s grid %s: %s ns\n" % (n, g, v_) for _, n, g, v_ in ls))
    open(os.path.join(P, "r02_cfg1_api_latency.txt"), "w").write(rd("r2c_api_latency.txt"))
    c = json.load(open(os.path.join(P, "r02_configs.json")))
    new = {}
    for f in ("r2c_cfg1.jsonl", "r2c_cfg3.jsonl", "r2c_xs.jsonl"):
        r = json.loads(rd(f).strip().splitlines()[-1]); new[r["config"]] = r
    c["records"] = [new.get(r.get("config"), r) for r in c["records"]]
    json.dump(c, open(os.path.join(P, "r02_configs.json"), "w"), indent=1)
    print("profiles/ updated")


if __name__ == "__main__":
    main()
