#!/bin/bash
# ncu launch list of reconstruct_scene (MATCH + ALIGN) on 8 views 512 x 512 (two passes; the second is the warm one)
set -u
TAG=${1:-r02}
export PYTHONPATH=$PWD
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/${TAG}_launches_reconstruct.csv python scripts/prof_reconstruct.py > gpurun_out/${TAG}_prof_reconstruct.log 2>&1
tail -3 gpurun_out/${TAG}_prof_reconstruct.log
