# First 2-GPU trip of a session (≈4 min wall, charged 2×): the multi-GPU suite incl. the tests added blind at the end of
# round 1 (stress test, padded SyncBN statistics, world-size dependent auto algorithm), then both bench arms.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/trip_first_2gpu.sh'
bash tools/trip_tests.sh 2
timeout -s KILL 200 python tools/numerics_probe.py > gpurun_out/numerics_probe.log 2>&1; grep -v "^syncbn" gpurun_out/numerics_probe.log | cut -c1-400 | tail -n 8
timeout -s KILL 200 python tools/allreduce_sweep.py --gpus 2 --max-mb 64 --out gpurun_out/sweep_2.json > gpurun_out/sweep_2.log 2>&1; tail -n 10 gpurun_out/sweep_2.log | cut -c1-300
P=$((29500 + RANDOM % 1000))
for cfg in "ours X=1" "ours_syncbn X=1 --syncbn" "ours_dbuf PDT_E2E_DOUBLE_BUFFER=1"; do
  set -- $cfg; name=$1; envv=$2; shift 2; P=$((P + 50))
  timeout -s KILL 200 env $envv python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 200 --warmup 20 "$@" > gpurun_out/bench_${name}_2.json 2> gpurun_out/bench_${name}_2.err
  cut -c1-330 gpurun_out/bench_${name}_2.json
done
tail -n 3 gpurun_out/bench_*_2.err | cut -c1-300
