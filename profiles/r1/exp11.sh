DPFT_NVCC_EXTRA="-DDPFT_DEBUG_STAMPS" python -c "
from deep_prob_feature_track_b200 import _lib
_lib.build(force=True)"
python profiles/exp11.py
