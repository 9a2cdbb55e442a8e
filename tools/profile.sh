#!/bin/bash
# ncu evidence for the round (run under gpurun, ONE GPU). Plain run first, ncu only if it exits 0.
#   tools/profile.sh <tag> [depth]
set -u
TAG=${1:-r01}
DEPTH=${2:-20}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --depth $DEPTH --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_hash -s 42 -c 2 -o $OUT/${TAG}_k_hash -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
tail -2 $OUT/${TAG}_plain.log
