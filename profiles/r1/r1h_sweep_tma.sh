#!/bin/bash
# TMA ring staging (tensor-map copies, -DDPFT_STAGED_BULK=2; 1 = per-map 1-D bulk copies) vs the cp.async ring
run() { for s in 1 8; do timeout 120 python bench.py --no-cpu-baseline --no-extras --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 | streams $s', round(d['value']), round(d['ms_per_step'],4), [round(x*1e3,1) for x in d['roofline']['all_launch_ms']][-6:])"; done; }
for mode in 1; do
  export DPFT_NVCC_EXTRA="-DDPFT_STAGED_TMA=$mode"
  python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" || continue
  echo "=== tma $mode"
  timeout 240 python -m pytest tests/test_uic_forward_gpu.py tests/test_edge_cases_gpu.py tests/test_keyframe_gpu.py -x -q 2>&1 | tail -4
  run "tma $mode"
done
