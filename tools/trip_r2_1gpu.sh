# Round-2 1-GPU trip: kernel tests (incl. the cooperative fused layers), bench with and without them.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/trip_r2_1gpu.sh'
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/kernel_tests.log 2>&1; tail -n 30 gpurun_out/kernel_tests.log | cut -c1-300
timeout -s KILL 240 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
cut -c1-2500 gpurun_out/bench_ours_1.json; tail -n 5 gpurun_out/bench_ours_1.err | cut -c1-400
PDT_FUSED_LAYERS=0 timeout -s KILL 240 python bench.py --steps 200 --warmup 20 --skip-e2e > gpurun_out/bench_ours_1_perop.json 2> gpurun_out/bench_ours_1_perop.err
cut -c1-400 gpurun_out/bench_ours_1_perop.json; tail -n 3 gpurun_out/bench_ours_1_perop.err | cut -c1-300
