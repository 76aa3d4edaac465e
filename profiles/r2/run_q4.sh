# round 2c: resident workers per SM of the one-map work-queue kernel; window loads with ld.global.cg
for v in q3w4cg q4w4cg q2w7r144cg; do
  echo "== $v"; DPFT_LIB_PATH=profiles/r2/variants/$v.so python profiles/r2/sigma_detect_probe.py 20 2>&1 | grep -E "passed in|Error|error"
done
v=q2w7r144cg
DPFT_LIB_PATH=profiles/r2/variants/$v.so ncu --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --section SpeedOfLight --clock-control none -k regex:uic_queue -s 2 -c 1 -o gpurun_out/diag_$v -f python profiles/r2/prof_target_queue.py 8 > gpurun_out/diag_$v.log 2>&1
