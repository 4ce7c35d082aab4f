mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi.txt 2>&1; nvidia-smi topo -m >> gpurun_out/smi.txt 2>&1
python tools/gpu_check.py > gpurun_out/check_stdout.log 2>&1
timeout 400 python bench.py --impl reference --steps 100 --warmup 10 > gpurun_out/bench_ref_1.json 2> gpurun_out/bench_ref_1.err
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ours_1.json 2> gpurun_out/bench_ours_1.err
timeout 300 python bench.py --steps 200 --warmup 20 --conv-impl simt > gpurun_out/bench_ours_simt.json 2> gpurun_out/bench_ours_simt.err
timeout 300 python bench.py --steps 100 --warmup 10 --no-graph --skip-e2e > gpurun_out/bench_ours_eager.json 2> gpurun_out/bench_ours_eager.err
cat gpurun_out/check_stdout.log; tail -3 gpurun_out/bench_*.err; cat gpurun_out/bench_*.json
