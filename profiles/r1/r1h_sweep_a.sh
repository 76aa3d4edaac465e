#!/bin/bash
# variants of the staged kernel: compile-time flags, quick parity check, bench at 1 and 8 streams
run() { for s in 1 8; do python bench.py --no-cpu-baseline --no-extras --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 | streams $s', round(d['value']), round(d['ms_per_step'],4), [round(x*1e3,1) for x in d['roofline']['all_launch_ms']])"; done; }
for flags in "" "-DDPFT_STAGED_ONE_BODY" "-DDPFT_GATHER_GROUP=8" "-DDPFT_GATHER_GROUP=2 -DDPFT_STAGED_CTAS=4 -DDPFT_STAGE_WIDTH=40" "-DDPFT_GATHER_GROUP=4 -DDPFT_STAGED_CTAS=4 -DDPFT_STAGE_WIDTH=40"; do
  export DPFT_NVCC_EXTRA="$flags"
  python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" || continue
  echo "=== [$flags]"
  python -m pytest tests/test_uic_forward_gpu.py tests/test_edge_cases_gpu.py -x -q 2>&1 | tail -1
  run "[$flags]"
done
