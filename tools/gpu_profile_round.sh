#!/bin/bash
# One profiling visit: (1) the ncu launch list of the bench command (durations only), (2) ncu --set full of every kernel of one
# MulRelinNew (k = 8, one lane) and of one hoisted Rotate, reduced to text on the box (gpurun_out carries at most 64 MiB back; only
# the capture of the dominant kernel keeps its .ncu-rep with sources).  usage: tools/gpu_profile_round.sh <tag>
T=${1:-r02q}; OUT=gpurun_out/$T; mkdir -p $OUT
if [ "${SKIP_LIST:-0}" != 1 ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
fi
# 3 warm-up steps x 2 ops x 20 launches = 120 launches (+ 2 k_tile_twiddles at context creation): skip them, capture one whole op
timeout 1500 ncu --set full --clock-control none -k regex:k_ -s 124 -c 21 -f -o /tmp/prof_mulrelin \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --batch 2 --lanes 1 > $OUT/ncu_mulrelin.log 2>&1
python tools/ncu_brief.py /tmp/prof_mulrelin.ncu-rep > $OUT/ncu_brief_mulrelin.txt 2>&1
ncu -i /tmp/prof_mulrelin.ncu-rep --page raw --csv > $OUT/ncu_raw_mulrelin.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:"k_mac_intt|k_moddown" -s 60 -c 3 -f -o /tmp/prof_rotate \
    python tools/rot_profile.py 4 > $OUT/ncu_rotate.log 2>&1
python tools/ncu_brief.py /tmp/prof_rotate.ncu-rep > $OUT/ncu_brief_rotate.txt 2>&1
ncu -i /tmp/prof_rotate.ncu-rep --page raw --csv > $OUT/ncu_raw_rotate.csv 2>/dev/null
# the dominant kernel with sources (hoisting launch of pass 2: the 4th launch of an op)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass2 -s 25 -c 1 -f -o $OUT/prof_pass2 \
    python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --batch 2 --lanes 1 > $OUT/ncu_pass2.log 2>&1
du -sh $OUT; ls -la $OUT; cat $OUT/ncu_brief_mulrelin.txt | head -60
