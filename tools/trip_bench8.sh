mkdir -p gpurun_out
N=${1:-8}
P=$((29500 + RANDOM % 1000))
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 300 --warmup 20 > gpurun_out/bench_ours_$N.json 2> gpurun_out/bench_ours_$N.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_ours_{N}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step")}, "verify", (d.get("verify") or {}).get("ok"), "e2e", e.get("ms_per_step"), e.get("value"))
    print(json.dumps(e.get("windows")))
except Exception as ex:
    print("no bench result:", ex)
PY
tail -n 2 gpurun_out/bench_ours_$N.err | cut -c1-300
