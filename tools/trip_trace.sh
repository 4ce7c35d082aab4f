mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/kernel_tests.log 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/kernel_tests.log | cut -c1-300 | tail -n 12
timeout -s KILL 120 python tools/fused_trace.py > gpurun_out/fused_trace.log 2>&1; cat gpurun_out/fused_trace.log | tail -n 45
timeout -s KILL 240 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours_1.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step", "windows")}, d.get("verify", {}).get("ok"))
    print("e2e", d.get("e2e"))
except Exception as e:
    print("no bench result:", e)
PY
tail -n 3 gpurun_out/bench_ours_1.err | cut -c1-300
