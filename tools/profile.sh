#!/bin/bash
# ncu evidence for the round (run under gpurun, ONE GPU). Plain run first, ncu only if it exits 0.
#   tools/profile.sh <tag> [depth]
# k_hash-matching launches per build: 1 leaf kernel + (depth - 14) one-thread-per-hash levels + 14 cooperative levels.
set -u
TAG=${1:-r01}
DEPTH=${2:-24}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --depth $DEPTH --steps 2 --warmup 3 --no-cpu-baseline"
PER_BUILD=$((DEPTH + 1))
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# leaf kernel + the first (largest) node level of the third build
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_hash -s $((2 * PER_BUILD)) -c 2 -o $OUT/${TAG}_k_hash -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
# the cooperative (3 lanes per hash) kernel on its largest level (8192 nodes) and on a 64-node level
$CMD > $OUT/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_hash2_coop -s 14 -c 8 -o $OUT/${TAG}_k_coop -f $CMD > $OUT/${TAG}_ncu_coop.log 2>&1
echo "coop capture rc=$?"
tail -1 $OUT/${TAG}_plain.log
