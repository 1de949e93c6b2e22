#!/bin/bash
# development: short profiled bench of the in-tree library and of variants built by tools/build_variant.sh.
# usage: tools/gpu_ab.sh <tag> [variant names...]
TAG=${1:-ab}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
show() { python - "$1" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("value", round(d["value"],1), d["unit"], "ms/op", round(d["ms_per_op"],4), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
print("   "+"  ".join(f"{k[2:]}:{v['ms_per_step']:.3f}" for k,v in d["kernels"].items()))
PY
}
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_default.json 2> $OUT/bench_default.err || tail -3 $OUT/bench_default.err
echo "== default"; show $OUT/bench_default.json
for v in "$@"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --lib tools/_build/libmkhe_$v.so > $OUT/bench_$v.json 2> $OUT/bench_$v.err || tail -3 $OUT/bench_$v.err
  echo "== $v"; show $OUT/bench_$v.json
done
