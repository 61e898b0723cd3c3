#!/bin/bash
# Run on the GPU box (under gpurun): plain run, ncu launch list, full captures of the two main kernels.
# Outputs land in gpurun_out/ and are summarised locally by tools/summarise_profiles.py into profiles/.
set -u
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/prof_ncu_list.log 2>&1
$CMD > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:voigt_tile -s 1 -c 1 -o gpurun_out/prof_voigt $CMD > gpurun_out/prof_ncu_voigt.log 2>&1
$CMD > gpurun_out/prof_plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tud_kernel -s 1 -c 1 -o gpurun_out/prof_tud $CMD > gpurun_out/prof_ncu_tud.log 2>&1
ls -la gpurun_out | tail -12
