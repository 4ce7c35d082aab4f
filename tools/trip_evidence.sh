# Evidence trip (1 GPU): per-kernel times vs the library ops, launch list, full ncu captures, sanitizers on the fused layers.
#   /usr/local/graft/bin/gpurun --timeout 1100 -- 'bash tools/trip_evidence.sh'
mkdir -p gpurun_out
timeout -s KILL 300 python tools/op_bench.py > gpurun_out/op_bench.log 2>&1; tail -n 32 gpurun_out/op_bench.log | cut -c1-200
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 8 -c 16 --csv --log-file gpurun_out/launches_fused.csv python tools/one_step.py 3 > gpurun_out/ncu_launches.log 2>&1
for k in convnet_fwd_kernel convnet_l2_bwd_kernel convnet_l1_bwd_kernel; do
  timeout -s KILL 240 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$k python tools/one_step.py 3 > gpurun_out/ncu_$k.log 2>&1
  tail -n 1 gpurun_out/ncu_$k.log | cut -c1-200
done
for tool in memcheck racecheck; do
  timeout -s KILL 300 compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 120 python tools/one_step.py 2 > gpurun_out/sanitizer_fused_$tool.log 2>&1
  echo "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|losses" gpurun_out/sanitizer_fused_$tool.log | cut -c1-300
done
ls -la gpurun_out/*.ncu-rep | tail -n 8
