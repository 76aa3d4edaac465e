DPFT_NVCC_EXTRA="-DDPFT_DEBUG_STAMPS" python -c "
from deep_prob_feature_track_b200 import _lib
_lib.build(force=True)"
python -m pytest tests/test_uic_forward_gpu.py -x -q 2>&1 | tail -2
python profiles/exp5.py 64 | tail -2
python profiles/exp5.py 8 | tail -2
