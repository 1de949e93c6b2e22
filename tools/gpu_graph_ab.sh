#!/bin/bash
# development: bench.py with the CUDA-graph cache on / off, interleaved repetitions on one box
T=${1:-ab}; mkdir -p gpurun_out/$T
for rep in 1 2 3; do
for mode in graphs nographs; do
  if [ $mode = nographs ]; then export MKHE_NO_GRAPHS=1; else unset MKHE_NO_GRAPHS; fi
  timeout 600 python bench.py --steps 20 --warmup 4 --no-cpu-baseline > gpurun_out/$T/bench_${mode}_$rep.json 2> gpurun_out/$T/bench_${mode}_$rep.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/$T/bench_${mode}_$rep.json").read().strip().splitlines()[-1])
e=d["extra"]
print("%-8s value %.1f e2e %.1f api %.1f rot8 %.0f mul4 %.1f rot4 %.0f c1 %.0f bfv %.1f cnn %.1f/%.1f clk %s %s" % ("$mode", d["value"], d["e2e"]["value"], e["mulrelin_new_reference_api_k8_ops_s"], e["rotate_hoisted_k8_ops_s"], e["mulrelin_k4_ops_s"], e["rotate_hoisted_k4_ops_s"], e["mulrelin_PN14QP439_k2_ops_s"], e["bfv_mulrelin_PN15QP880_k4_ops_s"], e["cnn_PN14QP433_images_s"], e["cnn_PN14QP433_e2e_images_s"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
PY
done; done
