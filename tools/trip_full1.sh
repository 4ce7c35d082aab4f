# 1-GPU: the whole GPU suite the way the driver runs it, smoke(), bench.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/gpu_suite_1.log 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/gpu_suite_1.log | cut -c1-300 | tail -n 12
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 | cut -c1-400
timeout -s KILL 240 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours_1.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step", "gpu_launches")}, "verify", (d.get("verify") or {}).get("ok"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("value"))
    print(d.get("clocks"))
except Exception as e:
    print("no bench result:", e)
PY
tail -n 3 gpurun_out/bench_ours_1.err | cut -c1-300
