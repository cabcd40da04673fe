#!/bin/bash
# Builds A/B variants of the tcgen05 matcher: starst3r_b200/libst3r_nn_<name>.so, name=flags pairs as arguments, e.g.
#   bash scripts/build_nn_variants.sh ser=-DNN_TC_EPI_SERIAL "dec=-DNN_TC_DECOUPLE"
set -e
cd "$(dirname "$0")/.."
python -m starst3r_b200.build > /dev/null
B=starst3r_b200/build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
OBJS=$(ls $B/*.o | grep -v "/nn_tc.o" | grep -v _var_)
for spec in "$@"; do
  name=${spec%%=*}; D=${spec#*=}
  nvcc $FLAGS $D -c starst3r_b200/csrc/nn_tc.cu -o $B/nn_tc_var_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o starst3r_b200/libst3r_nn_$name.so $OBJS $B/nn_tc_var_$name.o -cudart static -lpthread -ldl -lrt
done
ls -la starst3r_b200/*.so
