# round 2d: does leaving a third of every SM to the other stream's coarse levels pay?  (work-queue launch with 2 CTAs per SM)
for cfg in "2 0" "2 296" "3 296" "2 370" "3 0"; do set -- $cfg
  python bench.py --steps 64 --warmup 8 --streams $1 --queue-ctas $2 --no-cpu-baseline --no-extras --no-parity > gpurun_out/tmp_b.json 2> gpurun_out/tmp_b.err || tail -3 gpurun_out/tmp_b.err
  python -c "
import json
d=json.load(open('gpurun_out/tmp_b.json')); print('streams $1 queue_ctas $2:', round(d['value']), 'pairs/s', round(d['ms_per_step'],4), 'ms/step')"
done
