# single-GPU profiling trip: launch list of one eager step + full ncu captures of the top kernels
mkdir -p gpurun_out
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log | cut -c1-400
B="python bench.py --no-graph --steps 4 --warmup 3 --skip-e2e"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 160 --csv --log-file gpurun_out/launches_eager.csv $B > gpurun_out/ncu_eager.log 2>&1
for k in conv5x5_umma_tma_kernel conv5x5_wgrad_umma_kernel conv5x5_kernel linear_bwd_kernel bn_relu_pool_bwd_kernel; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -f -o gpurun_out/prof_$k $B > gpurun_out/ncu_$k.log 2>&1
  tail -n 2 gpurun_out/ncu_$k.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
# sanitizers on the single-GPU kernels (smoke = every ConvNet kernel fwd+bwd+step, 3 iterations)
for tool in memcheck racecheck; do
  timeout -s KILL 400 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke\]" gpurun_out/sanitizer_$tool.log | cut -c1-300
done
