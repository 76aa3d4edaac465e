# round 2d: row loop of the staged routine unrolled by two (independent rows for the scheduler to interleave)
for v in unroll2 unroll2c2; do
  echo "== $v"; DPFT_LIB_PATH=profiles/r2/variants/$v.so python profiles/r2/sigma_detect_probe.py 20 2>&1 | grep -E "passed in|Error|error"
done
