#!/bin/bash
# async-gather experiment: parity, then bench with and without it
python -m pytest tests/test_uic_forward_gpu.py -x -q -k "async" 2>&1 | tail -2
for extra in "" "--async-gather"; do
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$extra pairs/s %.0f  ms/step %.3f  lvl0 %.1f us frac %.3f  launches(us) %s' % (d['value'], d['ms_per_step'], r['launch_ms']*1e3, r['frac'], [round(x*1e3) for x in r['all_launch_ms']]))"
done
