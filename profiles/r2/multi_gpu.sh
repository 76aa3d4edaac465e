#!/bin/bash
# Round 2: multi-GPU runs of the three workloads (one rank per GPU, torchrun).  Usage (on an 8-GPU box):
#   bash profiles/r2/multi_gpu.sh            -> gpurun_out/r2_{tum,vga,train}_{N}gpu.json
cd "$(dirname "$0")/../.."
run() {  # workload N steps warmup extra...
  wl=$1; n=$2; steps=$3; warm=$4; shift 4
  if [ "$n" = 1 ]; then launcher="python"; else launcher="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n))"; fi
  timeout 600 $launcher bench.py --gpus $n --workload $wl --steps $steps --warmup $warm --no-cpu-baseline "$@" \
    > gpurun_out/r2_${wl}_${n}gpu.json 2> gpurun_out/r2_${wl}_${n}gpu.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_${wl}_${n}gpu.json"))
    e = d.get("e2e") or {}
    print("${wl} N=${n}: value %.1f pairs/s, ms/step %.3f, e2e %s, extra %s" % (d["value"], d["ms_per_step"], e.get("value"), json.dumps(d.get("extra", {}))[:300]))
except Exception as ex:
    print("${wl} N=${n}: FAILED", ex)
PY
  tail -2 gpurun_out/r2_${wl}_${n}gpu.err
}
for n in ${GPUS:-8 4 2}; do
  run train $n 10 3
  run vga $n 10 3 --no-parity
  run tum $n 20 5 --no-parity --no-extras
done
