#!/bin/bash
# Round-2 profile pass (run under gpurun): ncu launch list of the bench command, full captures of the Voigt kernel for the
# whole config-2 launch and per layer group, and of the layer-integration kernel.  Reports are summarised ON the box
# (tools/ncu_regions.py, tools/ncu_summary.py) and only the whole-launch reports travel back (gpurun_out/ <= 64 MiB).
set -u
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-banded"
O=gpurun_out
$CMD > $O/r2_prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2_launches.csv $CMD > $O/r2_prof_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:voigt_tile -s 1 -c 1 -f -o $O/r2_prof_voigt $CMD > $O/r2_prof_ncu_voigt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tud_kernel -s 1 -c 1 -f -o $O/r2_prof_tud $CMD > $O/r2_prof_ncu_tud.log 2>&1
python tools/ncu_summary.py $O/r2_prof_voigt.ncu-rep > $O/r2_voigt_summary.txt 2>&1
python tools/ncu_summary.py $O/r2_prof_tud.ncu-rep > $O/r2_tud_summary.txt 2>&1
python tools/ncu_regions.py $O/r2_prof_voigt.ncu-rep > $O/r2_regimes_all.txt 2>&1
for g in "0 10 l00" "10 30 l10" "30 50 l30" "50 66 l50"; do
  set -- $g
  python tools/layer_group_run.py $1 $2 > $O/r2_lg_$3.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:voigt_tile -s 1 -c 1 -f -o /tmp/r2_prof_voigt_$3 python tools/layer_group_run.py $1 $2 > $O/r2_lg_ncu_$3.log 2>&1
  python tools/ncu_regions.py /tmp/r2_prof_voigt_$3.ncu-rep > $O/r2_regimes_$3.txt 2>&1
done
ls -la $O | grep r2_
