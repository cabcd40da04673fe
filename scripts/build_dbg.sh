#!/bin/bash
# Instrumented variants of the library (development aid for the tcgen05 matcher):
#   libstarst3r_b200_dbg.so        cycle counters in the epilogue (st3r_debug_nn_tc_cycles)
#   libstarst3r_b200_nold.so       + no TMEM read, no arg-max arithmetic (timing only: the TMA + MMA pipeline alone)
#   libstarst3r_b200_split2.so     two epilogue warps per TMEM lane quarter (A/B against the default of one)
#   libstarst3r_b200_oneissuer.so  the first structure: ONE MMA thread for both query tiles, shared accumulator barriers
#                                  (-DNN_TC_ONE_ISSUER; the default has one issuing warp per query tile)
#   libstarst3r_b200_decouple.so   that thread with accumulator barriers per (query tile, stage): + -DNN_TC_DECOUPLE
#                                  (results identical; A/B with ST3R_B200_LIB=... python scripts/nn_variants.py)
# Select one with ST3R_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
python -m starst3r_b200.build > /dev/null
B=starst3r_b200/build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr"
OBJS=$(ls $B/*.o | grep -v nn_tc.o | grep -v _var_)
for v in dbg nold bn128 bn32 oneissuer decouple; do
  D=""
  N="-DNN_TC_DEBUG_CYCLES -DNN_TC_EXP_NOLD -DNN_TC_EXP_NOALU"
  [ $v = dbg ] && D="-DNN_TC_DEBUG_CYCLES"
  [ $v = nold ] && D="$N"
  [ $v = bn128 ] && D="-DNN_TC_BN=128"
  [ $v = bn32 ] && D="-DNN_TC_BN=32"
  [ $v = oneissuer ] && D="-DNN_TC_ONE_ISSUER"
  [ $v = decouple ] && D="-DNN_TC_ONE_ISSUER -DNN_TC_DECOUPLE"
  nvcc $FLAGS $D -c starst3r_b200/csrc/nn_tc.cu -o $B/nn_tc_var_$v.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o starst3r_b200/libstarst3r_b200_$v.so $OBJS $B/nn_tc_var_$v.o -cudart static -lpthread -ldl -lrt
done
# libstarst3r_b200_aligntime.so: per-phase cycle counts of the persistent ALIGN kernel (printed by the kernel)
OBJS2=$(ls $B/*.o | grep -v "/align.o" | grep -v _var_)
nvcc $FLAGS -DALIGN_PERSIST_TIMING -c starst3r_b200/csrc/align.cu -o $B/align_var_time.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o starst3r_b200/libstarst3r_b200_aligntime.so $OBJS2 $B/align_var_time.o -cudart static -lpthread -ldl -lrt
ls -la starst3r_b200/*.so
