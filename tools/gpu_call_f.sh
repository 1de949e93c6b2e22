#!/bin/bash
# development: multi-GPU visit; $1 = tag, $2 = number of GPUs
T=${1:-r02g}; N=${2:-2}
mkdir -p gpurun_out/$T
if [ "$N" = 1 ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$T/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$T/pytest.log; tail -3 gpurun_out/$T/pytest.log
  timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/$T/bench_n1.json 2> gpurun_out/$T/bench_n1.err; echo rc=$?
else
  timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/$T/bench_n$N.json 2> gpurun_out/$T/bench_n$N.err; echo rc=$?
  grep -v "NCCL INFO" gpurun_out/$T/bench_n$N.err | tail -5
fi
python - <<PY
import json
d=json.loads(open("gpurun_out/$T/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N value %.1f e2e %.1f ms/op %.4f scaling %s" % (d["value"], d["e2e"]["value"], d["ms_per_op"], d["scaling"]), (d.get("parity_check") or {}).get("compared"), {k:v for k,v in (d.get("sharded_parity") or {}).items() if k in ("equal_on_every_rank","timed_out")})
print("   e2e bytes", d["e2e"])
for k,v in d["kernels"].items(): print("   %-20s %.4f ms x%.0f hbm %.2f int %.2f" % (k, v["ms_per_step"], v["launches_per_step"], v.get("hbm_frac",0), v.get("int_frac",0)))
print(d["extra"])
PY
