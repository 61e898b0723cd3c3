#!/bin/bash
# Final records of round 2 (run under gpurun, one B200): bench line, launch list of the bench command, config 1 tools.
set -u
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r2c_bench_1gpu.json 2> $O/r2c_bench_1gpu.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-banded"
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/r2c_launches.csv $CMD > $O/r2c_ncu_list.log 2>&1
python tools/cfg1_run.py > $O/r2c_cfg1_tiles.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2c_cfg1_launches.csv python tools/cfg1_run.py > /dev/null 2>&1
python tools/api_latency.py 300 > $O/r2c_api_latency.txt 2>&1
python tools/run_configs.py 1 > $O/r2c_cfg1.jsonl 2>/dev/null
python tools/run_configs.py xs > $O/r2c_xs.jsonl 2>/dev/null
python tools/run_configs.py 3 > $O/r2c_cfg3.jsonl 2>/dev/null
ls -la $O | grep r2c_
