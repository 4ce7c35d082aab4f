# N-GPU bench trip without the test suite: bash tools/trip_n.sh N [ref] [sweep] [tests-k-expr:...]
N=$1; shift
mkdir -p gpurun_out
P=$((29500 + RANDOM % 1000))
for mode in "" "--syncbn"; do
  tag=ours${mode:+_syncbn}_$N
  timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 200 --warmup 20 $mode > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  P=$((P + 13))
  python - $tag <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{tag}.json").read().strip().splitlines()[-1])
    print(tag, {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step")}, "verify", (d.get("verify") or {}).get("ok"), (d.get("verify") or {}).get("grad_max_rel_err"),
          "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("clocks"))
except Exception as e:
    print(tag, "no bench result:", e)
PY
  tail -n 2 gpurun_out/bench_$tag.err | cut -c1-300
done
for extra in "$@"; do
  if [ "$extra" = ref ]; then
    timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P + 57)) bench.py --impl reference --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_ref_$N.json 2> gpurun_out/bench_ref_$N.err
    python - $N <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_ref_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print("ref", {k: d.get(k) for k in ("value", "ms_per_step")}, "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("ref: no result:", e)
PY
  elif [ "$extra" = sweep ]; then
    timeout -s KILL 300 python tools/allreduce_sweep.py --gpus $N --max-mb 64 --iters 100 --out gpurun_out/sweep_$N.json > gpurun_out/sweep_$N.log 2>&1; tail -n 26 gpurun_out/sweep_$N.log | cut -c1-220
  else
    PDT_TEST_WORLD=$N timeout -s KILL 400 python -m pytest tests/test_gpu_multigpu.py -q -m gpu --timeout 200 -p no:cacheprovider -k "$extra" > gpurun_out/comm_tests_$N.log 2>&1; grep -n "^E  \|passed\|failed" gpurun_out/comm_tests_$N.log | cut -c1-300 | tail -n 8
  fi
done
