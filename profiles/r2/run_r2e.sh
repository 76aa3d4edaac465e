# round 2e: bench lines of the build with up to three levels on the work queue (nothing here but the two ncu commands runs
# under a profiler; each ncu command follows a plain run of the same target that exited 0)
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_tum_1gpu_20steps.json 2> gpurun_out/r2e_tum_1gpu_20steps.err || tail -5 gpurun_out/r2e_tum_1gpu_20steps.err
python bench.py --workload vga > gpurun_out/r2e_vga_1gpu.json 2> gpurun_out/r2e_vga_1gpu.err || tail -5 gpurun_out/r2e_vga_1gpu.err
python bench.py --workload vga --frames-per-step 64 > gpurun_out/r2e_vga_1gpu_64frames.json 2> gpurun_out/r2e_vga_1gpu_64frames.err || tail -5 gpurun_out/r2e_vga_1gpu_64frames.err
python profiles/r2/launch_list_target.py 20 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_call20.csv python profiles/r2/launch_list_target.py 20 > gpurun_out/r2e_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uic_queue_kernel -c 8 -o gpurun_out/r2e_queue_levels -f python profiles/r2/launch_list_target.py 20 > gpurun_out/r2e_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
