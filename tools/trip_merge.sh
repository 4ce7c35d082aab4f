# 1-GPU trip: kernel tests, phase trace, bench with conv2's weight gradient riding on layer-1 backward vs as its own kernel.
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/kernel_tests.log 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/kernel_tests.log | cut -c1-300 | tail -n 12
timeout -s KILL 120 python tools/fused_trace.py > gpurun_out/fused_trace.log 2>&1; cat gpurun_out/fused_trace.log | tail -n 50
for m in ${MODES:-1 0}; do
  PDT_WGRAD_MERGED=$m timeout -s KILL 240 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1_m$m.json 2> gpurun_out/bench_ours_1_m$m.err
  python - $m <<'PY'
import json, sys
m = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_ours_1_m{m}.json").read().strip().splitlines()[-1])
    print("merged=" + m, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step")}, d.get("windows", {}).get("median_ms_per_step"), d.get("verify", {}).get("ok"), "e2e", d.get("e2e", {}).get("ms_per_step"))
except Exception as e:
    print("no bench result:", e)
PY
  tail -n 2 gpurun_out/bench_ours_1_m$m.err | cut -c1-300
done
cp gpurun_out/bench_ours_1_m1.json gpurun_out/bench_ours_1.json
