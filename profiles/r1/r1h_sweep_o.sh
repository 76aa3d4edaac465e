#!/bin/bash
# 480x640 workload: parked outlier lanes per row (4 vs 8), and a 2-way wait policy at 120x160
runv() { python bench.py --workload vga --no-cpu-baseline --no-extras --streams 8 --steps 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 | vga', round(d['value']), round(d['ms_per_step'],4), [round(x*1e3,1) for x in d['roofline']['all_launch_ms']][-3:])"; }
run() { python bench.py --no-cpu-baseline --no-extras --streams 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 | tum', round(d['value']), round(d['ms_per_step'],4), [round(x*1e3,1) for x in d['roofline']['all_launch_ms']][-3:])"; }
runv "out lanes 4"; run "out lanes 4"
export DPFT_NVCC_EXTRA="-DDPFT_OUT_LANES=8"
python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" || exit 1
python -m pytest tests/test_edge_cases_gpu.py -x -q 2>&1 | tail -1
runv "out lanes 8"; run "out lanes 8"
