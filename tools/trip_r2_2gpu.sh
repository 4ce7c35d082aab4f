# Round-2 2-GPU trip: multi-GPU suite, both bench arms at N=2, SyncBN numerics probe.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/trip_r2_2gpu.sh'
mkdir -p gpurun_out
bash tools/trip_tests.sh 2 "$1"
P=$((29500 + RANDOM % 1000))
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_ours_2.json 2> gpurun_out/bench_ours_2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_ours_2.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches_per_step", "windows", "verify", "details")})
    print("e2e", {k: v for k, v in (d.get("e2e") or {}).items() if k != "note"})
except Exception as e:
    print("no bench result:", e)
PY
tail -n 5 gpurun_out/bench_ours_2.err | cut -c1-400
P=$((P + 57))
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --impl reference --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_ref_2.json 2> gpurun_out/bench_ref_2.err
cut -c1-900 gpurun_out/bench_ref_2.json; tail -n 3 gpurun_out/bench_ref_2.err | cut -c1-300
grep -c "_C.so\|pytorch_distributed_train_b200" /dev/null
timeout -s KILL 200 python tools/numerics_probe.py > gpurun_out/numerics_probe.log 2>&1; grep -i "resnet\|syncbn" gpurun_out/numerics_probe.log | cut -c1-300 | tail -n 14
