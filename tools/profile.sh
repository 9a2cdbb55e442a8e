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
if [ "${PROFILE_BUILD:-1}" = 1 ]; then
# launch list of THIS library's kernels only (-k regex:^k_): bench.py synthesises its leaves with a few hundred tiny torch
# element-wise launches during (untimed) setup, which would otherwise fill the capture window. 6 builds x (depth+1)
# launches covers the setup build, the warm-up and the timed steps; shares are per build, so any whole build will do.
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c $((6 * PER_BUILD)) --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# leaf kernel + the first (largest) node level of the third build
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_hash -s $((2 * PER_BUILD)) -c 2 -o $OUT/${TAG}_k_hash -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
fi
# the cooperative (3 lanes per hash) kernel on its largest level (8192 nodes) and on a 64-node level
if [ "${PROFILE_COOP:-1}" = 1 ]; then
$CMD > $OUT/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_hash_coop -s 14 -c 8 -o $OUT/${TAG}_k_coop -f $CMD > $OUT/${TAG}_ncu_coop.log 2>&1
echo "coop capture rc=$?"
fi
# the witness-trace fold (BASELINE config 4): 2^14 paths of the depth-20 tree = 4.2 GB of trace per launch
if [ "${PROFILE_TRACE:-1}" = 1 ]; then
TCMD="python bench.py --workload paths --depth 20 --queries 16384 --steps 1 --warmup 1"
$TCMD > $OUT/${TAG}_plain_trace.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fold_paths -c 1 -o $OUT/${TAG}_k_trace -f $TCMD > $OUT/${TAG}_ncu_trace.log 2>&1
echo "trace capture rc=$?"
fi
# the same traces read from the resident tree: one independent hash per (query, level)
if [ "${PROFILE_TREE_TRACE:-1}" = 1 ]; then
TCMD="python bench.py --workload paths --depth 20 --queries 16384 --steps 1 --warmup 1"
$TCMD > $OUT/${TAG}_plain_tree_trace.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_tree_paths -c 1 -o $OUT/${TAG}_k_tree_trace -f $TCMD > $OUT/${TAG}_ncu_tree_trace.log 2>&1
echo "tree trace capture rc=$?"
fi
tail -1 $OUT/${TAG}_plain.log
