mkdir -p gpurun_out
python tools/gpu_check.py ddp1 > gpurun_out/check_stdout3.log 2>&1
# per-kernel device times of eager steps (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 140 --csv --log-file gpurun_out/launches_eager.csv python bench.py --steps 6 --warmup 3 --no-graph --skip-e2e > gpurun_out/ncu_eager.log 2>&1
# full capture of the tcgen05 conv kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv5x5_umma -s 4 -c 2 -o gpurun_out/prof_conv_umma -f python bench.py --steps 4 --warmup 3 --no-graph --skip-e2e > gpurun_out/ncu_conv.log 2>&1
cat gpurun_out/check_stdout3.log; tail -n 3 gpurun_out/ncu_eager.log | cut -c1-200; ls -la gpurun_out/
