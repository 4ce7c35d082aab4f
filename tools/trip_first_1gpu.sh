# First 1-GPU trip of a session (≈4–5 min): everything that was prepared blind at the end of round 1.
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/trip_first_1gpu.sh'
mkdir -p gpurun_out
# 1. does the tree still pass on hardware?  (kernel tests only; the multi-GPU file needs ≥2 GPUs)
timeout -s KILL 400 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider > gpurun_out/kernel_tests.log 2>&1; tail -n 3 gpurun_out/kernel_tests.log
# 2. hardware probes for the window conv
timeout -s KILL 120 python tools/exp_rowshift.py > gpurun_out/exp_rowshift.log 2>&1; cat gpurun_out/exp_rowshift.log | cut -c1-400
for m in 1 0; do PDT_WIN_BASE_OFFSET=$m timeout -s KILL 120 python tools/check_conv_win.py > gpurun_out/check_conv_win_bo$m.log 2>&1; echo "base_offset=$m"; tail -n 5 gpurun_out/check_conv_win_bo$m.log | cut -c1-300; done
# 3. headline bench, and the experimental knobs one at a time
timeout -s KILL 200 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
PDT_E2E_DOUBLE_BUFFER=1 timeout -s KILL 200 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1_dbuf.json 2> gpurun_out/bench_ours_1_dbuf.err
PDT_CONV_IMPL=win timeout -s KILL 200 python bench.py --steps 200 --warmup 20 --skip-e2e > gpurun_out/bench_ours_1_win.json 2> gpurun_out/bench_ours_1_win.err
for f in gpurun_out/bench_ours_1*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"), d.get("clocks"))
except Exception as e:
    print("no result:", e)
PY
done
tail -n 3 gpurun_out/bench_ours_1*.err | cut -c1-300
# 4. profiles and sanitizers
bash tools/trip_ncu.sh
timeout -s KILL 200 python tools/op_bench.py > gpurun_out/op_bench.log 2>&1; tail -n 30 gpurun_out/op_bench.log
