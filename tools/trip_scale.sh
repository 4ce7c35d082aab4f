# scaling trip: usage  bash tools/trip_scale.sh <N>   (benches first, then sweep, ResNet-18, tests)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
# preflight: bring the heap + multicast up at this world size with a short leash; fall back to P2P-only if it fails
if ! timeout -s KILL 150 python tools/allreduce_sweep.py --gpus $N --max-mb 1 --iters 20 > gpurun_out/preflight_$N.log 2>&1; then
  echo "PREFLIGHT FAILED with multicast; retrying without"; tail -n 5 gpurun_out/preflight_$N.log | cut -c1-300
  export PDT_SYMM_NO_MULTICAST=1
  timeout -s KILL 150 python tools/allreduce_sweep.py --gpus $N --max-mb 1 --iters 20 > gpurun_out/preflight_nomc_$N.log 2>&1 || { echo "PREFLIGHT FAILED again"; tail -n 5 gpurun_out/preflight_nomc_$N.log | cut -c1-300; exit 1; }
fi
tail -n 2 gpurun_out/preflight_$N.log | cut -c1-300
P=$((29500 + RANDOM % 1000))
run() { name=$1; shift; P=$((P + 50))
  timeout -s KILL 300 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N $ARGS > gpurun_out/bench_${name}_$N.json 2> gpurun_out/bench_${name}_$N.err
}
ARGS="--steps 200 --warmup 20" run ours X=1
ARGS="--impl reference --steps 60 --warmup 10" run ref X=1
ARGS="--steps 200 --warmup 20 --syncbn" run ours_syncbn X=1
ARGS="--impl reference --steps 60 --warmup 10 --syncbn --skip-e2e" run ref_syncbn X=1
cat gpurun_out/bench_*_$N.json | cut -c1-400
timeout -s KILL 400 python tools/allreduce_sweep.py --gpus $N --max-mb 1024 --out gpurun_out/sweep_$N.json > gpurun_out/sweep_$N.log 2>&1
tail -n 14 gpurun_out/sweep_$N.log | cut -c1-330
for impl in ours reference; do
  P=$((P + 50))
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/bench_resnet.py --impl $impl --steps 20 --warmup 5 > gpurun_out/resnet_${impl}_$N.json 2> gpurun_out/resnet_${impl}_$N.err
done
cat gpurun_out/resnet_*_$N.json | cut -c1-600; tail -n 3 gpurun_out/resnet_*_$N.err | cut -c1-300
export PDT_TEST_WORLD=$N
timeout -s KILL 600 python -m pytest tests/test_gpu_multigpu.py -q -m gpu --timeout 200 -p no:cacheprovider -k "collectives or nccl or fused or lockstep or absent" > gpurun_out/comm_tests_$N.log 2>&1
grep -n "^E  \|passed\|failed" gpurun_out/comm_tests_$N.log | cut -c1-300 | tail -n 30
tail -n 3 gpurun_out/bench_*_$N.err | cut -c1-300
