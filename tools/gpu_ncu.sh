#!/bin/bash
# one ncu --set full capture of a few kernels inside a short bench run.  usage: tools/gpu_ncu.sh <tag> <kernel-regex> [skip] [count]
TAG=${1:-n}; RE=${2:-k_ntt_pass2}; SKIP=${3:-8}; CNT=${4:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$RE" -s $SKIP -c $CNT -f -o $OUT/prof \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --batch 2 > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log; ls -la $OUT
