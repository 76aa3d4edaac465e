python -m pytest tests/test_sigma_detect_gpu.py tests/test_uic_queue_gpu.py tests/test_uic_forward_gpu.py tests/test_keyframe_gpu.py tests/test_edge_cases_gpu.py -x -q -m gpu 2>&1 | tail -3
python profiles/r2/sigma_detect_probe.py 20 2>&1 | grep -E "default|passed in|Error|error"
