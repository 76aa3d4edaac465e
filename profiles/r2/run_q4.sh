# round 2d: deterministic fold of the ICP term's sums; pixels per thread of icp_term_kernel (its 27-sum epilogue is amortised over them)
python -m pytest tests/test_uic_forward_gpu.py tests/test_uic_backward_gpu.py tests/test_context_gpu.py tests/test_reference_dropin_gpu.py -x -q -m gpu 2>&1 | tail -2
python profiles/r2/icp_launch_target.py 2>&1 | tail -2
for v in icp_t1 icp_t05; do echo "== $v"; DPFT_LIB_PATH=profiles/r2/variants/$v.so python profiles/r2/icp_launch_target.py 2>&1 | tail -2 | head -1; done
