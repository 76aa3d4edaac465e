#!/bin/bash
run() { for s in 1 8; do python bench.py --no-cpu-baseline --no-extras --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 | streams $s', round(d['value']), round(d['ms_per_step'],4), [round(x*1e3,1) for x in d['roofline']['all_launch_ms']])"; done; }
export DPFT_NVCC_EXTRA="-DDPFT_DEBUG_STAMPS"
python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" || exit 1
echo "=== dealt tiling, timeline"
python profiles/r1h_timeline.py 2>&1 | grep -v "^   (\|  seg"
export DPFT_NVCC_EXTRA=""
python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" || exit 1
python -m pytest tests/test_uic_forward_gpu.py tests/test_edge_cases_gpu.py tests/test_keyframe_gpu.py -x -q 2>&1 | tail -1
run "dealt"
DPFT_LINEAR_TILES=1 run "linear"
DPFT_CTA_SLOTS=444 run "dealt slots=444"
