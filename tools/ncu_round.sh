#!/bin/bash
# the round's ncu evidence for the bench workload (run on the GPU box): launch list of two suite passes, one --set full capture
# of one pass, both exported as CSV (the reports themselves do not fit gpurun_out/)
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT; TMP=/tmp/ncu_round; mkdir -p $TMP
CMD="python bench.py --steps 4 --warmup 1 --suite-passes 1 --streams 1 --no-timeline --device-only"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_ncu_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --launch-skip 37 --launch-count 37 -f -o $TMP/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
ncu -i $TMP/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_full_raw.csv 2>> $OUT/${TAG}_ncu_full.log
ls -la $OUT/${TAG}_ncu_*
