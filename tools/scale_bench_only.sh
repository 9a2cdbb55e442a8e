mkdir -p gpurun_out
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_bench_gpus$N.json 2> gpurun_out/r02f_bench_gpus$N.err; echo "bench $N rc=$?"
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/r02f_bench_gpus1.json 2> gpurun_out/r02f_bench_gpus1.err; echo "bench 1 rc=$?"
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    d = json.loads(open(f"gpurun_out/r02f_bench_gpus{n}.json").read().strip().splitlines()[-1])
    base = base or d["value"]
    print(f"N={n} value {d['value'] / 1e6:.2f} M hashes/s  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value'] / 1e6:.2f} M  efficiency {d['value'] / base / n:.3f}  leaf kernel {d['roofline'].get('kernel_ms_per_launch', 0):.2f} ms  node levels {d['roofline'].get('node_levels_ms_per_step', 0):.2f} ms  root {d['root'][:12]}  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
PY
