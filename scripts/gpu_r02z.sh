#!/bin/bash
# GPU pass r02z: verify the reworked persistent ALIGN loop (pair-straddling rows walked per pair), parity suite, bench, ncu of the loop.
set -u
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== align breakdown, variant 7 / 3"
for v in 7 3; do
  ST3R_ALIGN_VARIANT=$v timeout 300 python scripts/prof_align.py 2>&1 | tail -1 | tee $OUT/${TAG}_align_breakdown_v$v.json
done
echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["step_breakdown_ms"])
print(d["reconstruct"]["seconds"], d["reconstruct"]["stages_s"], d["match"]["value"], d["reconstruct"].get("align"))
PY
tail -3 $OUT/${TAG}_bench.err
echo "== ncu align_persist"
ST3R_PROF_PASSES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:align_persist_kernel -c 1 -f \
    -o $OUT/${TAG}_prof_align_persist python scripts/prof_align.py > $OUT/ncu_align_persist.log 2>&1
tail -2 $OUT/ncu_align_persist.log
python scripts/ncu_summary.py $OUT/${TAG}_prof_align_persist.ncu-rep | head -30 | tee $OUT/${TAG}_ncu_summary.txt
