#!/bin/bash
# One 8-GPU box, one call:   gpurun --gpus 8 -- tools/scale_check.sh [tag]
#   1. tools/multi_gpu_check.sh 8   (plain-C clients in both modes, the Python layer under torchrun, tests/test_gpu_multi.py, bench --gpus 8)
#   2. the torchrun check + bench at N = 4 on the first 4 GPUs
#   3. bench at N = 2 and N = 1: the 1 -> 8 scaling table of ONE box (profiles/<tag>_scaling.md is written from these lines)
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
bash tools/multi_gpu_check.sh 8 $TAG > $OUT/${TAG}_scale_check.log 2>&1
LOG=$OUT/${TAG}_multi_gpu_check_4.log
echo "== torchrun tools/multi_gpu_check.py (4 GPUs)" > $LOG
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/multi_gpu_check.py 16 >> $LOG 2>&1; echo "rc=$?" >> $LOG
tests/_build/cabi_multi_driver multi 4 16 >> $LOG 2>&1; echo "rc=$?" >> $LOG
for N in 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_gpus$N.json 2>> $LOG; echo "bench $N rc=$?" >> $LOG
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_bench_gpus1.json 2>> $LOG; echo "bench 1 rc=$?" >> $LOG
python - <<'PY' $TAG
import json, sys
tag = sys.argv[1]
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"gpurun_out/{tag}_bench_gpus{n}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "missing", e); continue
    base = base or d["value"]
    sec = d.get("secondary", {}).get("paths", {})
    print(f"N={n} value {d['value'] / 1e6:.2f} M hashes/s  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value'] / 1e6:.2f} M  efficiency {d['value'] / base / n:.3f}  "
          f"leaf kernel {d['roofline'].get('kernel_ms_per_launch', 0):.2f} ms  root {d['root'][:12]}  paths {sec.get('value', 0) / 1e6:.1f} M traced hashes/s e2e {sec.get('e2e', {}).get('value', 0) / 1e6:.1f} M")
PY
grep -v "^\[W\|^W0\|^\*\*\*\|Setting OMP" $OUT/${TAG}_scale_check.log | grep "rc=\|ok\|== \|sharded\|passed\|failed" | head -30
cat $LOG | grep "rc=\|ok" | head
