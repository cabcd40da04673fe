#!/bin/bash
set -u
export PYTHONPATH=$PWD
for lib in starst3r_b200/libst3r_var_dbg*.so; do
  echo "== $lib"
  ST3R_B200_LIB=$PWD/$lib timeout 300 python scripts/dbg_tc_cycles.py 2>&1 | tail -2
done
