#!/bin/bash
# channel-split sweep: DPFT_SPLIT values given as arguments
python -m pytest tests/test_uic_forward_gpu.py -x -q 2>&1 | tail -1
for sp in "$@"; do
  export DPFT_SPLIT="$sp"
  echo "=== DPFT_SPLIT=$sp"
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('pairs/s %.0f  ms/step %.3f  lvl0 launch %.1f us  frac %.3f  launches(us) %s' % (d['value'], d['ms_per_step'], r['launch_ms']*1e3, r['frac'], [round(x*1e3) for x in r['all_launch_ms']]))"
done
