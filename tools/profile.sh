#!/bin/bash
# ncu evidence for the round (run under gpurun, ONE GPU). Every capture follows a plain run of the same command that exited 0.
#   tools/profile.sh <tag> [depth]       then, in the container:  python tools/ncu_summary.py <tag>
# Launches of one build at depth >= 16 (two half-trees on two streams): 1 leaf kernel + 2 (depth - 1) half-levels + 1 root level.
set -u
TAG=${1:-r02}
DEPTH=${2:-24}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --depth $DEPTH --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
PER_BUILD=$((2 * DEPTH))
# thread-per-hash launches of one build: the leaf kernel + the half-levels above 4096 nodes per half
TPH=$((1 + 2 * (DEPTH - 14)))
# imt_ctx_create's self-test launches each kernel family first: 2 x k_hash, 2 x k_hash_coop, 2 x k_hash_lh
SELF=2
# latency-kernel launches of one build: 3 lanes per hash for the half-levels of 4096 / 2048 / 1024 nodes per half (6), the
# lead / helper kernel for the 10 smaller half-levels + the root (21)
COOP=6
LH=21
if [ "${PROFILE_BUILD:-1}" = 1 ]; then
# launch list of THIS library's kernels only (-k regex:^k_): bench.py synthesises its leaves with a few hundred tiny torch
# element-wise launches during (untimed) setup, which would otherwise fill the capture window. Shares are per build.
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c $((6 * PER_BUILD)) --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# leaf kernel + the first (largest) half-level of the third build
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:^k_hash$ -s $((SELF + 2 * TPH)) -c 2 -o $OUT/${TAG}_k_hash -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
fi
# the 3-lanes-per-hash latency kernel: the 4096-nodes-per-half level and smaller ones of one build
if [ "${PROFILE_COOP:-1}" = 1 ]; then
$CMD > $OUT/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_hash_coop -s $((SELF + COOP)) -c 6 -o $OUT/${TAG}_k_coop -f $CMD > $OUT/${TAG}_ncu_coop.log 2>&1
echo "coop capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_hash_lh -s $((SELF + LH)) -c 3 -o $OUT/${TAG}_k_lh -f $CMD > $OUT/${TAG}_ncu_lh.log 2>&1
echo "lead/helper capture rc=$?"
fi
# witness traces read from the resident tree: one independent traced hash per (query, level)
if [ "${PROFILE_TREE_TRACE:-1}" = 1 ]; then
TCMD="python bench.py --workload paths --depth 20 --queries 16384 --steps 1 --warmup 1"
$TCMD > $OUT/${TAG}_plain_tree_trace.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_trace_tree_paths -c 1 -o $OUT/${TAG}_k_tree_trace -f $TCMD > $OUT/${TAG}_ncu_tree_trace.log 2>&1
echo "tree trace capture rc=$?"
fi
# the low-leaf lookup: 2^20 queries against the depth-24 index (prefix array + shared-memory top), and the plain binary search
if [ "${PROFILE_LOOKUP:-1}" = 1 ]; then
LCMD="python bench.py --workload lookups --depth $DEPTH --steps 1 --warmup 1"
$LCMD > $OUT/${TAG}_plain_lookup.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_low_leaf_lookup -c 1 -o $OUT/${TAG}_k_lookup -f $LCMD > $OUT/${TAG}_ncu_lookup.log 2>&1
echo "lookup capture rc=$?"
IMT_FAST_LOOKUP_MIN=1000000000 ncu --set full --clock-control none --import-source on -k regex:k_low_leaf_lookup -c 1 -o $OUT/${TAG}_k_lookup_plain -f $LCMD > $OUT/${TAG}_ncu_lookup_plain.log 2>&1
echo "plain lookup capture rc=$?"
fi
tail -1 $OUT/${TAG}_plain.log | cut -c1-400
