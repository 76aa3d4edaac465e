# round 2d: accumulate_system with fp32x2 pairs (-DDPFT_PACKED_ACCUM)
DPFT_LIB_PATH=profiles/r2/variants/pacc.so python -m pytest tests/test_uic_queue_gpu.py tests/test_uic_forward_gpu.py tests/test_edge_cases_gpu.py tests/test_sigma_detect_gpu.py tests/test_keyframe_gpu.py -x -q -m gpu 2>&1 | tail -2
echo "== packed accumulate"; DPFT_LIB_PATH=profiles/r2/variants/pacc.so python profiles/r2/sigma_detect_probe.py 20 2>&1 | grep -E "default|check off  |passed in|Error|error"
echo "== default (with the extreme pre-test)"; python profiles/r2/sigma_detect_probe.py 20 2>&1 | grep -E "default|check off  |passed in|Error|error"
