# multi-GPU trip: usage  bash tools/trip_mg.sh <N> [quick]
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
export PDT_TEST_WORLD=$N
timeout -s KILL 900 python -m pytest tests/test_gpu_multigpu.py -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/comm_tests_$N.log 2>&1
tail -n 15 gpurun_out/comm_tests_$N.log | cut -c1-300
timeout -s KILL 400 python tools/allreduce_sweep.py --gpus $N --max-mb 64 --out gpurun_out/sweep_$N.json > gpurun_out/sweep_$N.log 2>&1
tail -n 12 gpurun_out/sweep_$N.log | cut -c1-400
P=$((29500 + RANDOM % 1000))
run() { # name, extra env/args...
  name=$1; shift
  P=$((P + 50))
  timeout -s KILL 400 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N $ARGS > gpurun_out/bench_${name}_$N.json 2> gpurun_out/bench_${name}_$N.err
}
ARGS="--impl reference --steps 60 --warmup 10" run ref X=1
ARGS="--steps 200 --warmup 20" run ours X=1
ARGS="--steps 200 --warmup 20 --skip-e2e" run ours_unfused PDT_FUSE_OPT=0
ARGS="--steps 200 --warmup 20 --syncbn" run ours_syncbn X=1
tail -n 4 gpurun_out/bench_*_$N.err | cut -c1-300; cat gpurun_out/bench_*_$N.json | cut -c1-1300
