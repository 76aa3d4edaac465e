#!/bin/bash
# one tuning experiment: parity subset, bench (both launch modes), dynamic instruction count of a level-0 launch
python -m pytest tests/test_uic_forward_gpu.py -x -q -k "golden or full_size_vs_oracle or single_level" 2>&1 | tail -1
for extra in "" "--single-launch"; do
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$extra pairs/s %.0f  ms/step %.3f  lvl0 %.1f us frac %.3f  launches(us) %s' % (d['value'], d['ms_per_step'], r['launch_ms']*1e3, r['frac'], [round(x*1e3) for x in r['all_launch_ms']]))"
done
python profiles/prof_target2.py > /dev/null 2>&1 && ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:uic_iter_kernel -s 21 -c 1 --csv python profiles/prof_target2.py 2>/dev/null | grep -E "inst_executed|time_duration|issue_active" | awk -F'","' '{print $(NF-2), $NF}'
