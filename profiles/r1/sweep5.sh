#!/bin/bash
for flags in "$@"; do
  export DPFT_NVCC_EXTRA="$flags"
  python -c "from deep_prob_feature_track_b200 import _lib; _lib.build(force=True)" || exit 1
  echo "=== $flags"
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras  2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('pairs/s %.0f  ms/step %.3f  lvl0 launch %.1f us  frac %.3f  launches(us) %s' % (d['value'], d['ms_per_step'], r['launch_ms']*1e3, r['frac'], [round(x*1e3) for x in r['all_launch_ms']]))"
done
