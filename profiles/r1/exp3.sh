run() { python bench.py --no-cpu-baseline --no-extras --streams 1 --async-gather 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['roofline']['all_launch_ms'])"; }
python -c "
from deep_prob_feature_track_b200 import _lib
_lib.build(force=True)"
run base
DPFT_NVCC_EXTRA="-DDPFT_EXPERIMENT_NO_SOBEL" python -c "
from deep_prob_feature_track_b200 import _lib
_lib.build(force=True)"
run nosobel
