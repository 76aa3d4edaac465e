# round 2d: final 1-GPU measurements (tests first; nothing here runs under a profiler)
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r2d_gpu_tests.log 2>&1; tail -2 gpurun_out/r2d_gpu_tests.log
python bench.py > gpurun_out/r2d_tum_1gpu.json 2> gpurun_out/r2d_tum_1gpu.err || tail -5 gpurun_out/r2d_tum_1gpu.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_tum_1gpu_20steps.json 2> gpurun_out/r2d_tum_1gpu_20steps.err || tail -5 gpurun_out/r2d_tum_1gpu_20steps.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_reference_arm.json 2> gpurun_out/r2d_reference_arm.err || tail -5 gpurun_out/r2d_reference_arm.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
