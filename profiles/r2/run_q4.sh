python -m pytest tests/test_uic_queue_gpu.py -x -q -m gpu -k graph_replay 2>&1 | tail -30
DPFT_LIB_PATH=profiles/r2/variants/kr_g1.so python -m pytest tests/test_uic_queue_gpu.py -x -q -m gpu -k graph_replay 2>&1 | tail -30
v=kr_g1
DPFT_LIB_PATH=profiles/r2/variants/$v.so ncu --set full --import-source on --clock-control none -k regex:uic_queue -s 2 -c 1 -o gpurun_out/prof_$v -f python profiles/r2/prof_target_queue.py 8 > gpurun_out/prof_$v.log 2>&1
