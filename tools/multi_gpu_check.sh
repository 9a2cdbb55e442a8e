#!/bin/bash
# Multi-GPU checks of the C-ABI path, run on an N-GPU box:   gpurun --gpus N -- tools/multi_gpu_check.sh N [tag]
#   1. plain C, ONE process, N devices  (imt_multi_create -> ncclCommInitAll): sharded depth-16 tree == single-GPU tree == golden root
#   2. plain C, one process PER GPU     (imt_comm_create  -> ncclCommInitRank, id through a file): the same, no Python anywhere
#   3. the Python layer over the same calls under torchrun (tools/multi_gpu_check.py): paths, lookups, traces, inserts vs 1 GPU + oracle
#   4. tests/test_gpu_multi.py (now over NCCL instead of the copy transport)
#   5. bench.py --gpus N (short)
N=${1:-2}
TAG=${2:-r02}
OUT=gpurun_out
mkdir -p $OUT
LOG=$OUT/${TAG}_multi_gpu_check_$N.log
: > $LOG
EXE=tests/_build/cabi_multi_driver
if [ ! -x $EXE ]; then
  mkdir -p tests/_build
  gcc -std=c99 -O1 -Iinclude tests/cabi_multi_driver.c -o $EXE -Lindexed-merkle-tree-halo2_b200 -limt_b200 -Wl,-rpath,$PWD/indexed-merkle-tree-halo2_b200 >> $LOG 2>&1
fi
echo "== 1. C client, one process, $N devices" >> $LOG
NCCL_DEBUG=WARN $EXE multi $N 16 >> $LOG 2>&1; echo "rc=$?" >> $LOG
echo "== 2. C client, one process per GPU" >> $LOG
rm -f /tmp/imt_nccl.id
for r in $(seq 0 $((N-1))); do NCCL_DEBUG=WARN $EXE rank $r $N /tmp/imt_nccl.id 16 > $OUT/${TAG}_crank_$r.log 2>&1 & done
wait
for r in $(seq 0 $((N-1))); do echo "-- rank $r" >> $LOG; cat $OUT/${TAG}_crank_$r.log >> $LOG; done
echo "== 3. torchrun tools/multi_gpu_check.py" >> $LOG
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py 16 >> $LOG 2>&1; echo "rc=$?" >> $LOG
echo "== 4. pytest tests/test_gpu_multi.py" >> $LOG
python -m pytest tests/test_gpu_multi.py -m gpu -x -q >> $LOG 2>&1; echo "rc=$?" >> $LOG
echo "== 5. bench.py --gpus $N" >> $LOG
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_gpus$N.json 2>> $LOG; echo "rc=$?" >> $LOG
tail -c 1500 $OUT/${TAG}_bench_gpus$N.json >> $LOG
grep -v "^\[W\|^W0\|^\*\*\*\|Setting OMP" $LOG | tail -60
