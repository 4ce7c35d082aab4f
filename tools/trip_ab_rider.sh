mkdir -p gpurun_out
for r in 1 0 1 0; do
  PDT_SGD_RIDER=$r timeout -s KILL 240 python bench.py --steps 400 --warmup 20 --skip-verify > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - $r <<'PY'
import json, sys
d = json.loads(open("gpurun_out/ab.json").read().strip().splitlines()[-1])
print("rider", sys.argv[1], round(d["ms_per_step"], 5), d["gpu_launches_per_step"], "e2e", round(d["e2e"]["ms_per_step"], 5), d["windows"]["median_ms_per_step"])
PY
done
timeout -s KILL 120 python tools/fused_trace.py 2>&1 | grep -A12 "replayed CUDA graph" | grep -A9 "l1_bwd"
