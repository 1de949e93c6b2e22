#!/bin/bash
# One gpurun visit: GPU parity tests, smoke, a short bench, the ncu launch list and one full capture of the
# top kernel.  Everything lands in gpurun_out/.   usage: tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
tail -3 $OUT/launches.csv
echo "== ncu full capture of ${NCU_KERNEL:-k_ntt_pass2}"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${NCU_KERNEL:-k_ntt_pass2} -s 4 -c 2 -f -o $OUT/prof_top \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
fi
