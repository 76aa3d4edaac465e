# round 2d: final 1-GPU measurements (tests first; ncu only after the plain runs have exited 0)
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2d_gpu_tests.log 2>&1; tail -2 gpurun_out/r2d_gpu_tests.log
python bench.py > gpurun_out/r2d_tum_1gpu.json 2> gpurun_out/r2d_tum_1gpu.err || tail -5 gpurun_out/r2d_tum_1gpu.err
python bench.py --workload vga > gpurun_out/r2d_vga_1gpu.json 2> gpurun_out/r2d_vga_1gpu.err || tail -5 gpurun_out/r2d_vga_1gpu.err
python bench.py --workload train > gpurun_out/r2d_train_1gpu.json 2> gpurun_out/r2d_train_1gpu.err || tail -5 gpurun_out/r2d_train_1gpu.err
ncu --set full --import-source on --clock-control none -k regex:uic_queue_kernel -s 4 -c 1 -o gpurun_out/prof_r2d_queue_onemap_G8 -f python profiles/r2/prof_target_queue.py 8 > gpurun_out/prof_r2d.log 2>&1; tail -2 gpurun_out/prof_r2d.log
