# N-GPU trip (N = $1): multi-GPU suite, our bench arm (default and --syncbn), optionally the reference arm and the numerics probe.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/trip_multi.sh 2 "" ref probe'
N=${1:-2}; K="$2"; shift 2
mkdir -p gpurun_out
bash tools/trip_tests.sh $N "$K"
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
          "exposed_us", (d.get("details") or {}).get("backward_comm_exposed_us"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("value"))
except Exception as e:
    print(tag, "no bench result:", e)
PY
  tail -n 3 gpurun_out/bench_$tag.err | cut -c1-300
done
for extra in "$@"; do
  if [ "$extra" = ref ]; then
    timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P + 57)) bench.py --impl reference --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_ref_$N.json 2> gpurun_out/bench_ref_$N.err
    cut -c1-400 gpurun_out/bench_ref_$N.json; tail -n 2 gpurun_out/bench_ref_$N.err | cut -c1-300
  elif [ "$extra" = probe ]; then
    timeout -s KILL 240 python tools/numerics_probe.py > gpurun_out/numerics_probe.log 2>&1; grep "^\[probe\]" gpurun_out/numerics_probe.log | cut -c1-420 | tail -n 24
  elif [ "$extra" = sweep ]; then
    timeout -s KILL 300 python tools/allreduce_sweep.py --gpus $N --out gpurun_out/sweep_$N.json > gpurun_out/sweep_$N.log 2>&1; tail -n 30 gpurun_out/sweep_$N.log | cut -c1-200
  fi
done
