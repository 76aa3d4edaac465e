#!/bin/bash
# round-1 (session h) final validation: GPU tests, both bench arms, launch list and one full capture of the level-0 launch
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_r1h_ref.json 2> gpurun_out/bench_r1h_ref.err
python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err
python bench.py --streams 1 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_r1h_1stream.json
python bench.py --streams 1 --no-pdl --no-cpu-baseline --no-extras 2>/dev/null | tail -1 > gpurun_out/bench_r1h_1stream_nopdl.json
python profiles/prof_target4.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1h.csv python profiles/prof_target4.py > gpurun_out/ncu_r1h_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:uic_iter_staged_kernel -s 9 -c 1 -o gpurun_out/prof_r1h_staged -f python profiles/prof_target4.py > gpurun_out/ncu_r1h.log 2>&1
python profiles/prof_target5.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1h_train.csv python profiles/prof_target5.py > gpurun_out/ncu_r1h_train.log 2>&1
tail -c 1500 gpurun_out/bench_r1h.json
