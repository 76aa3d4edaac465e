# round 2d: pixels per thread of icp_term_kernel (its 27-sum epilogue is amortised over them)
python profiles/r2/icp_launch_target.py 2>&1 | tail -2
for v in icp_t1 icp_t05 icp_t025; do echo "== $v"; DPFT_LIB_PATH=profiles/r2/variants/$v.so python profiles/r2/icp_launch_target.py 2>&1 | tail -2 | head -1; done
