for cfg in "4 4" "4 3" "2 5" "2 6"; do
set -- $cfg
DPFT_NVCC_EXTRA="-DDPFT_BWD_GROUP=$1 -DDPFT_BWD_CTAS=$2" python -c "
from deep_prob_feature_track_b200 import _lib
_lib.build(force=True)" || continue
python bench.py --no-cpu-baseline --steps 50 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('group $1 ctas $2', d['extra']['train_step']['ms_per_step'])"
done
