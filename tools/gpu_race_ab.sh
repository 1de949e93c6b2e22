#!/bin/bash
# development: the back-to-back stress on ONE box: the in-tree library (fence before every release of a TMA landing buffer) against
# a build without the fence (tools/build_variant.sh nofence -DMKHE_NO_RELEASE_FENCE), with and without the L2 eviction hints
R=${1:-400}
nvidia-smi --query-gpu=uuid,serial,pci.bus_id,temperature.gpu --format=csv,noheader
echo "== in-tree (fence), hints on";    timeout 600 python tools/stress_b2b.py $R 16 2 2>&1 | tail -3
echo "== in-tree (fence), hints off";   MKHE_DEBUG_NO_L2_HINTS=1 timeout 600 python tools/stress_b2b.py $R 16 2 2>&1 | tail -3
echo "== in-tree (fence), hints off, k=4";   MKHE_DEBUG_NO_L2_HINTS=1 timeout 600 python tools/stress_b2b.py $((R/4)) 16 4 2>&1 | tail -3
if [ -f tools/_build/libmkhe_nofence.so ]; then
echo "== no fence, hints off";          MKHE_LIB=tools/_build/libmkhe_nofence.so MKHE_DEBUG_NO_L2_HINTS=1 timeout 600 python tools/stress_b2b.py $R 16 2 2>&1 | tail -3
fi
