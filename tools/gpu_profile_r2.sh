#!/bin/bash
# Round-2 profile pass (run under gpurun): ncu launch list of the bench command, full captures of the Voigt kernel for the
# whole config-2 launch and per layer group, and of the layer-integration kernel.  Reports are summarised ON the box
# (tools/ncu_regions.py, tools/ncu_summary.py) and only the whole-launch reports travel back (gpurun_out/ <= 64 MiB).
set -u
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-banded"
O=gpurun_out
$CMD > $O/r2_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2_launches.csv $CMD > $O/r2_prof_ncu_list.log 2>&1
# the automatic tile size launches the K = 8 and the K = 4 grid (one retires at once): capture with the size pinned to
# what the device picks for config 2 (K = 8, --flags 2048)
ncu --set full --clock-control none --import-source on -k regex:voigt_tile -s 1 -c 1 -f -o $O/r2_prof_voigt $CMD --flags 2048 > $O/r2_prof_ncu_voigt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tud_kernel -s 1 -c 1 -f -o $O/r2_prof_tud $CMD > $O/r2_prof_ncu_tud.log 2>&1
python tools/ncu_summary.py $O/r2_prof_voigt.ncu-rep > $O/r2_voigt_summary.txt 2>&1
python tools/ncu_summary.py $O/r2_prof_tud.ncu-rep > $O/r2_tud_summary.txt 2>&1
python tools/ncu_regions.py $O/r2_prof_voigt.ncu-rep > $O/r2_regimes_all.txt 2>&1
for g in "0 10 l00 8" "10 30 l10 8" "30 50 l30 8" "50 66 l50 4"; do
  set -- $g
  python tools/layer_group_run.py $1 $2 2 $4 > $O/r2_lg_$3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:voigt_tile -s 1 -c 1 -f -o /tmp/r2_prof_voigt_$3 python tools/layer_group_run.py $1 $2 2 $4 > $O/r2_lg_ncu_$3.log 2>&1
  python tools/ncu_regions.py /tmp/r2_prof_voigt_$3.ncu-rep > $O/r2_regimes_$3.txt 2>&1
done
ls -la $O | grep r2_
# sensor kernels (HBM-bound): full captures, summarised on the box
for k in "ils ils_tma r02_ils_tma_ncu" "app apparent_kernel r02_apparent_ncu" "app32 apparent_kernel r02_apparent_f32_ncu" "hsi hsi_kernel r02_hsi_ncu" "hsi32 hsi_kernel r02_hsi_f32_ncu"; do
  set -- $k
  python tools/sensor_ncu_run.py $1 > /dev/null 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o /tmp/$3 python tools/sensor_ncu_run.py $1 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/$3.ncu-rep 0.03 > $O/$3.txt 2>&1
done
# the FP64 peak the roofline divides by: the dependent-free DFMA loop under ncu (pipe utilisation = evidence for the 34.7 TFLOP/s)
python -c "import torch; from radtxfr_b200 import engine; print(engine.default_handle('cuda:0').fp64_peak())" > $O/r02_dfma_peak_plain.log 2>&1
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max --clock-control none -k regex:dfma_peak -c 4 python -c "import torch; from radtxfr_b200 import engine; print(engine.default_handle('cuda:0').fp64_peak())" > $O/r02_dfma_peak_ncu.txt 2>&1
# sensor microbenchmarks and the other configurations
python tools/sensor_microbench.py > $O/r02_sensor_microbench.txt 2>&1
python tools/voigt_microbench.py 5 > $O/r02_voigt_microbench.txt 2>&1
python tools/run_configs.py 1 3 4 > $O/r02_configs_134.jsonl 2> $O/r02_configs_134.err
python tools/cfg1_run.py > $O/r02_cfg1_tiles.txt 2>&1
ls -la $O | grep -E "r2_|r02_"
