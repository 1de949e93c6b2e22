#!/bin/bash
# development: one GPU visit = tests + A/B bench + rotate profiles
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r02a/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a/pytest.log
tail -5 gpurun_out/r02a/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a/bench_new.json 2> gpurun_out/r02a/bench_new.err; echo "bench new rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lib build/libmkhe_b200_r1.so > gpurun_out/r02a/bench_r1.json 2> gpurun_out/r02a/bench_r1.err; echo "bench r1 rc=$?"
for k in 4 8; do
  timeout 120 python tools/rot_profile.py $k > gpurun_out/r02a/rot_new_k$k.txt 2>&1
  timeout 120 python tools/rot_profile.py $k build/libmkhe_b200_r1.so > gpurun_out/r02a/rot_r1_k$k.txt 2>&1
done
cat gpurun_out/r02a/rot_new_k4.txt gpurun_out/r02a/rot_r1_k4.txt
