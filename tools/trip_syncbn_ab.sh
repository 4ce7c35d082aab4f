mkdir -p gpurun_out
P=$((29500 + RANDOM % 1000))
for cfg in "PDT_LOADER_WORKERS=1 --syncbn" "PDT_LOADER_WORKERS=2 --syncbn" "PDT_LOADER_WORKERS=1" ; do
  set -- $cfg
  env $1 timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 200 --warmup 20 $2 > gpurun_out/ab.json 2> gpurun_out/ab.err
  P=$((P + 11))
  python - "$cfg" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}
    print(sys.argv[1], "device", round(d["ms_per_step"], 4), "e2e", round(e.get("ms_per_step", 0), 4), e.get("windows", {}).get("host_median_ms"))
except Exception as ex:
    print(sys.argv[1], "failed", ex)
PY
done
timeout -s KILL 240 python tools/numerics_probe.py > gpurun_out/numerics_probe.log 2>&1; grep "^\[probe\] resnet" gpurun_out/numerics_probe.log | cut -c1-330
