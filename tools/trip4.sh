mkdir -p gpurun_out
python tools/gpu_check.py conv_exact conv_bwd conv_fwd convnet ddp1 > gpurun_out/check_stdout4.log 2>&1
timeout -s KILL 300 python tools/op_bench.py > gpurun_out/op_bench.log 2>&1
timeout -s KILL 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_eager.csv python bench.py --steps 5 --warmup 3 --no-graph --skip-e2e > gpurun_out/ncu_eager.log 2>&1
cat gpurun_out/check_stdout4.log gpurun_out/op_bench.log; tail -n 3 gpurun_out/bench_ours_1.err; cat gpurun_out/bench_ours_1.json
