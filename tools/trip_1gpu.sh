mkdir -p gpurun_out
python tools/gpu_check.py loaded conv_exact conv_tma conv_fwd conv_bwd bn_pool head_sgd convnet ddp1 > gpurun_out/check_stdout.log 2>&1
timeout -s KILL 300 python tools/op_bench.py > gpurun_out/op_bench.log 2>&1
timeout -s KILL 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
cat gpurun_out/check_stdout.log gpurun_out/op_bench.log; tail -n 3 gpurun_out/bench_ours_1.err | cut -c1-300; cut -c1-1500 gpurun_out/bench_ours_1.json
