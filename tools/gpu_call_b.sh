#!/bin/bash
# development: quick GPU visit = selected tests + bench + rotate profiles; $1 = output tag
T=${1:-r02b}
mkdir -p gpurun_out/$T
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/$T/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$T/pytest.log
tail -4 gpurun_out/$T/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/$T/bench.json 2> gpurun_out/$T/bench.err; echo "bench rc=$?"
for k in 4 8; do timeout 120 python tools/rot_profile.py $k > gpurun_out/$T/rot_k$k.txt 2>&1; done
cat gpurun_out/$T/rot_k4.txt gpurun_out/$T/rot_k8.txt
python - <<PY
import json
d=json.loads(open("gpurun_out/$T/bench.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms/op %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_op"]), (d.get("parity_check") or {}).get("compared"))
for k,v in d["kernels"].items(): print("   %-20s %.4f ms x%.0f" % (k, v["ms_per_step"], v["launches_per_step"]))
print(d["extra"])
PY
