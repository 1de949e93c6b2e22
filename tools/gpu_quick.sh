#!/bin/bash
# quick GPU visit: parity tests + a short profiled bench (per-kernel CUDA-event times).  usage: tools/gpu_quick.sh [tag]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "clocks", d["clocks"])
for k,v in d["kernels"].items(): print(f"  {k:28s} {v['ms_per_step']:.3f} ms  share {v['share']:.3f}  hbm {v.get('hbm_frac',0):.2f} int {v.get('int_frac',0):.2f}")
print("extra", d.get("extra"))
PY
