#!/bin/bash
# development: L2 hint A/B on hoisted Rotate + ncu captures
T=${1:-r02c}
mkdir -p gpurun_out/$T
for k in 4 8; do
  timeout 120 python tools/rot_profile.py $k > gpurun_out/$T/rot_hint_k$k.txt 2>&1
  MKHE_DEBUG_NO_L2_HINTS=1 timeout 120 python tools/rot_profile.py $k > gpurun_out/$T/rot_nohint_k$k.txt 2>&1
done
head -6 gpurun_out/$T/rot_hint_k4.txt gpurun_out/$T/rot_nohint_k4.txt gpurun_out/$T/rot_hint_k8.txt gpurun_out/$T/rot_nohint_k8.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mac_intt|k_moddown' -s 12 -c 3 -o gpurun_out/$T/rot_hint -f python tools/rot_profile.py 4 > gpurun_out/$T/ncu_hint.log 2>&1
MKHE_DEBUG_NO_L2_HINTS=1 timeout 600 ncu --set full --clock-control none -k regex:'k_mac_intt' -s 4 -c 1 -o gpurun_out/$T/rot_nohint -f python tools/rot_profile.py 4 > gpurun_out/$T/ncu_nohint.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/$T/bench.json 2> gpurun_out/$T/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/$T/bench.json").read().strip().splitlines()[-1])
print("value %.1f e2e %.1f ms/op %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_op"]))
for k,v in d["kernels"].items(): print("   %-20s %.4f ms x%.0f" % (k, v["ms_per_step"], v["launches_per_step"]))
print(d["extra"])
PY
ls -la gpurun_out/$T
