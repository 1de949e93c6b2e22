#!/bin/bash
# development: GPU tests + bench with the graph cache on and off.  $1 = tag
T=${1:-r02o}
mkdir -p gpurun_out/$T
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/$T/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$T/pytest.log; tail -3 gpurun_out/$T/pytest.log
for mode in graphs nographs; do
  if [ $mode = nographs ]; then export MKHE_NO_GRAPHS=1; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$T/bench_$mode.json 2> gpurun_out/$T/bench_$mode.err; echo "$mode rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/$T/bench_$mode.json").read().strip().splitlines()[-1])
print("$mode: value %.1f e2e %.1f launches %s graphs %s" % (d["value"], d["e2e"]["value"], d["gpu_launches"], {k:v for k,v in d["cuda_graphs"].items() if k!="note"}))
print("   ", {k:(round(v,1) if isinstance(v,float) else v) for k,v in d["extra"].items() if not isinstance(v,dict)})
PY
done
