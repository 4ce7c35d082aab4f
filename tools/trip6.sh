mkdir -p gpurun_out
python tools/gpu_check.py conv_exact conv_tma conv_fwd conv_bwd convnet ddp1 > gpurun_out/check_stdout6.log 2>&1
timeout -s KILL 300 python tools/op_bench.py > gpurun_out/op_bench.log 2>&1
timeout -s KILL 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
PDT_WGRAD_TCGEN05=1 timeout -s KILL 300 python bench.py --steps 200 --warmup 20 --skip-e2e > gpurun_out/bench_ours_wg.json 2> gpurun_out/bench_ours_wg.err
cat gpurun_out/check_stdout6.log gpurun_out/op_bench.log; tail -n 3 gpurun_out/bench_ours_*.err | cut -c1-300; cut -c1-330 gpurun_out/bench_ours_1.json gpurun_out/bench_ours_wg.json
