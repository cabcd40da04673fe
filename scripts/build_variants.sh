#!/bin/bash
# Builds A/B variants of ONE source file of the library: starst3r_b200/libst3r_var_<name>.so.
#   bash scripts/build_variants.sh gs_raster.cu slots3072=-DST3R_POOL_SLOTS=3072 "pg128=-DST3R_POOL_PG=128 -DST3R_POOL_SLOTS=2560"
# Select one with ST3R_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
SRC=$1; shift
STEM=${SRC%.cu}
python -m starst3r_b200.build > /dev/null
B=starst3r_b200/build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
# the two files whose arithmetic must not be FMA-contracted (starst3r_b200/build.py NO_FMAD)
case $SRC in gs_project.cu|gs_backward.cu) FLAGS="$FLAGS -fmad=false";; esac
OBJS=$(ls $B/*.o | grep -v "/$STEM.o" | grep -v _var_)
for spec in "$@"; do
  name=${spec%%=*}; D=${spec#*=}
  nvcc $FLAGS $D -c starst3r_b200/csrc/$SRC -o $B/${STEM}_var_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o starst3r_b200/libst3r_var_$name.so $OBJS $B/${STEM}_var_$name.o -cudart static -lpthread -ldl -lrt
done
ls -la starst3r_b200/*.so
