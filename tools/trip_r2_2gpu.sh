# Round-2 2-GPU trip: multi-GPU suite (all tests, incl. the former "experimental" ones), ours bench at N=2 and N=1.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 900 -- 'bash tools/trip_r2_2gpu.sh'
mkdir -p gpurun_out
bash tools/trip_tests.sh 2 "$1"
P=$((29500 + RANDOM % 1000))
timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/bench_ours_2.json 2> gpurun_out/bench_ours_2.err
cut -c1-1800 gpurun_out/bench_ours_2.json; tail -n 5 gpurun_out/bench_ours_2.err | cut -c1-400
timeout -s KILL 240 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
cut -c1-1800 gpurun_out/bench_ours_1.json; tail -n 5 gpurun_out/bench_ours_1.err | cut -c1-400
