#!/bin/bash
# development: hoisting chunk (L2 residency) experiment + DRAM bytes per MulRelin
T=${1:-r02e}
mkdir -p gpurun_out/$T
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/$T/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/$T/pytest.log
tail -3 gpurun_out/$T/pytest.log
for hd in 0 2 4 7; do
  MKHE_DEBUG_HOIST_DIGITS=$hd timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/$T/bench_hd$hd.json 2> gpurun_out/$T/bench_hd$hd.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/$T/bench_hd$hd.json").read().strip().splitlines()[-1])
print("hoist_digits=$hd value %.1f e2e %.1f ms/op %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_op"]), {k: round(v["ms_per_step"],4) for k,v in d["kernels"].items() if k in ("k_ntt_pass2","k_bcast_ntt_pass1","k_mac_parties","k_mac_intt","k_moddown_Q")})
PY
done
for hd in 0 2 4; do
  MKHE_DEBUG_HOIST_DIGITS=$hd timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/$T/ncu_dram_hd$hd.csv python bench.py --steps 1 --warmup 3 --batch 2 --lanes 1 --no-extras --no-cpu-baseline > gpurun_out/$T/ncu_hd$hd.log 2>&1
done
for k in 4 8; do timeout 120 python tools/rot_profile.py $k > gpurun_out/$T/rot_k$k.txt 2>&1; done
cat gpurun_out/$T/rot_k4.txt gpurun_out/$T/rot_k8.txt
