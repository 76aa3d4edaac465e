#!/bin/bash
# re-check of the tiling / column / group options after the scoreboard fix
run() { for s in 1 8; do python bench.py --no-cpu-baseline --no-extras --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 | streams $s', round(d['value']), round(d['ms_per_step'],4), [round(x*1e3,1) for x in d['roofline']['all_launch_ms']][-6:])"; done; }
run "default"
DPFT_LINEAR_TILES=1 run "linear tiles"
DPFT_CTA_SLOTS=432 run "slots 432"
for flags in "-DDPFT_STAGED_COLS=32" "-DDPFT_GATHER_GROUP=8" "-DDPFT_STAGED_WARPS=2"; do
  export DPFT_NVCC_EXTRA="$flags"
  python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" || continue
  python -m pytest tests/test_uic_forward_gpu.py -x -q 2>&1 | tail -1
  run "[$flags]"
done
