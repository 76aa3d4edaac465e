#!/bin/bash
# round-1 (session h) final validation: GPU tests, both bench arms, stream-count check, launch lists
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r1h_ref.json 2> gpurun_out/bench_r1h_ref.err
python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err
for s in 1 6 12; do python bench.py --streams $s --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_r1h_${s}streams.json; done
python profiles/prof_target4.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv python profiles/prof_target4.py > gpurun_out/ncu_r1h_launches.log 2>&1
python profiles/prof_target5.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1h_train.csv python profiles/prof_target5.py > gpurun_out/ncu_r1h_train.log 2>&1
tail -c 600 gpurun_out/bench_r1h.json
# the TMA ring variant stays parity-green (compile-time option, off by default)
DPFT_NVCC_EXTRA="-DDPFT_STAGED_TMA=1" python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" && timeout 240 python -m pytest tests/test_uic_forward_gpu.py tests/test_edge_cases_gpu.py tests/test_keyframe_gpu.py -x -q 2>&1 | tail -2
