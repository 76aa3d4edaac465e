run() { for s in 1 8; do python bench.py --no-cpu-baseline --no-extras --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 streams $s', round(d['value']), d['ms_per_step'], d['roofline']['all_launch_ms'][-3:])"; done; }
for cfg in "3 5" "3 4" "6 2" "2 6"; do
set -- $cfg
DPFT_NVCC_EXTRA="-DDPFT_WARPS=$1 -DDPFT_STAGED_CTAS=$2" python -c "
from deep_prob_feature_track_b200 import _lib
_lib.build(force=True)" || continue
python -m pytest tests/test_uic_forward_gpu.py -x -q -k "staged and full_size" 2>&1 | tail -1
run "warps=$1 ctas=$2"
done
