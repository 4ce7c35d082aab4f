mkdir -p gpurun_out
python tools/gpu_check.py conv_fwd conv_bwd convnet ddp1 > gpurun_out/check_stdout2.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
timeout 300 python bench.py --steps 200 --warmup 20 --conv-impl simt > gpurun_out/bench_ours_simt.json 2> gpurun_out/bench_ours_simt.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-graph --skip-e2e > gpurun_out/bench_ours_eager.json 2> gpurun_out/bench_ours_eager.err
# per-kernel device times of the eager step (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 120 --csv --log-file gpurun_out/launches_eager.csv python bench.py --steps 8 --warmup 3 --no-graph --skip-e2e > gpurun_out/ncu_eager.log 2>&1
cat gpurun_out/check_stdout2.log; tail -3 gpurun_out/bench_ours*.err; cat gpurun_out/bench_ours*.json
