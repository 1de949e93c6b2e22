#!/bin/bash
# development: build a variant of the library with extra -D flags.  usage: tools/build_variant.sh <name> [-DFLAG=..]...
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p tools/_build
NCCL=$(python -c "import __graft_entry__ as g; print(' '.join(g._nccl_flags()))")
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Iinclude "$@" \
    mkhe_kklss_b200/csrc/mkhe_api.cu -o tools/_build/libmkhe_$NAME.so $NCCL
