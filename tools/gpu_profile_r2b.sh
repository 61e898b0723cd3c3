#!/bin/bash
# Round-2 (second session) profile pass, run under gpurun: launch list of the bench command, full captures of the Voigt
# kernel (whole config-2 launch, tile size and variant pinned: --flags 18432 = K 8 | RTX_VOIGT_NOTILEFAR, which is what
# the device picks for every state of config 2 except the top layers) and of the layer-integration kernel; summaries
# are made ON the box and only the two whole-launch reports travel back.
set -u
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-banded"
O=gpurun_out
$CMD > $O/r2b_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/r2b_launches.csv $CMD > $O/r2b_prof_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voigt_tile -s 1 -c 1 -f -o $O/r2b_prof_voigt $CMD --flags 18432 > $O/r2b_prof_ncu_voigt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tud_kernel -s 1 -c 1 -f -o $O/r2b_prof_tud $CMD > $O/r2b_prof_ncu_tud.log 2>&1
python tools/ncu_summary.py $O/r2b_prof_voigt.ncu-rep > $O/r2b_voigt_summary.txt 2>&1
python tools/ncu_summary.py $O/r2b_prof_tud.ncu-rep > $O/r2b_tud_summary.txt 2>&1
python tools/ncu_regions.py $O/r2b_prof_voigt.ncu-rep > $O/r2b_regimes_all.txt 2>&1
python tools/voigt_microbench.py 5 > $O/r2b_voigt_microbench.txt 2>&1
python tools/tilefar_sweep.py 3 > $O/r2b_tilefar_sweep.txt 2>&1
ls -la $O | grep r2b_
