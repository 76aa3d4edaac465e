#!/bin/bash
# balanced linear tiling with 2-warp CTAs vs 4-warp CTAs; slot-count sweep
run() { for s in 1 8; do python bench.py --no-cpu-baseline --no-extras --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 | streams $s', round(d['value']), round(d['ms_per_step'],4), [round(x*1e3,1) for x in d['roofline']['all_launch_ms']])"; done; }
for flags in "" "-DDPFT_STAGED_WARPS=4" "-DDPFT_STAGED_WARPS=1"; do
  export DPFT_NVCC_EXTRA="$flags"
  python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" || continue
  echo "=== [$flags]"
  python -m pytest tests/test_uic_forward_gpu.py tests/test_edge_cases_gpu.py tests/test_keyframe_gpu.py -x -q 2>&1 | tail -3
  run "[$flags]"
  if [ -z "$flags" ]; then
    for slots in 888 864 816; do DPFT_CTA_SLOTS=$slots run "[$flags] slots=$slots"; done
  fi
done
