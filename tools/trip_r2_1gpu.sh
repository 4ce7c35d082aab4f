# Round-2 1-GPU trip: kernel tests (incl. the cooperative fused layers), bench, per-launch device times.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/trip_r2_1gpu.sh'
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/kernel_tests.log 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/kernel_tests.log | cut -c1-300 | tail -n 25
timeout -s KILL 240 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours_1.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step", "windows", "verify")})
    print("e2e", d.get("e2e"))
except Exception as e:
    print("no bench result:", e)
PY
tail -n 5 gpurun_out/bench_ours_1.err | cut -c1-400
# per-launch device times of two eager steps (ncu serialises and flushes caches: compare shares, not absolutes)
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_eager.csv python bench.py --no-graph --steps 4 --warmup 3 --skip-e2e --skip-verify > gpurun_out/ncu_eager.log 2>&1
python - <<'PY'
import csv
try:
    rows = [r for r in csv.reader(open("gpurun_out/launches_eager.csv")) if len(r) > 10 and r[0].isdigit()]
    for r in rows[-20:]:
        print(f"{float(r[-1]):8.2f} us  {r[4][:110]}")
except Exception as e:
    print("no launch list:", e)
PY
