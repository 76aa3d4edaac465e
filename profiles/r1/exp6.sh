run() { for s in 1 8; do python bench.py --no-cpu-baseline --no-extras --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 streams $s', round(d['value']), d['ms_per_step'], d['roofline']['all_launch_ms'][-3:])"; done; }
python -m pytest tests/test_uic_forward_gpu.py -x -q -k staged 2>&1 | tail -1
run now
