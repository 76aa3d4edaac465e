# round 2c: robust (majority) footprint of the staged routine
python -m pytest tests/test_edge_cases_gpu.py tests/test_uic_forward_gpu.py tests/test_uic_queue_gpu.py tests/test_keyframe_gpu.py tests/test_sigma_detect_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== majority footprint"; python profiles/r2/vga_probe2.py 2>&1 | grep -E "motion|Error|error"
python profiles/r2/sigma_detect_probe.py 20 2>&1 | grep -E "default|passed in|Error|error"
echo "== minimum footprint (before)"; DPFT_LIB_PATH=profiles/r2/variants/norobust.so python profiles/r2/vga_probe2.py 2>&1 | grep -E "motion|Error|error"
DPFT_LIB_PATH=profiles/r2/variants/norobust.so python profiles/r2/sigma_detect_probe.py 20 2>&1 | grep -E "default|passed in|Error|error"
