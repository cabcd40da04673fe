#!/bin/bash
# GPU pass r02aa: CUDA-graph replay of the training iteration (gs.TRAIN_GRAPH): tests, bench with / without.
set -u
TAG=${1:-r02aa}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== pytest gs (graph tests first)"
timeout 600 python -m pytest tests/test_gs_gpu.py -m gpu -x -q -k "adam or graph or plan or scene_api or train_steps" 2>&1 | tail -8
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== bench (graph)"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["gpu_launches"], d.get("graph_replays"), d["roofline"]["step_breakdown_ms"])
PY
tail -3 $OUT/${TAG}_bench.err
echo "== bench (eager, train leg only)"
ST3R_TRAIN_GRAPH=0 timeout 900 python bench.py > $OUT/${TAG}_bench_eager.json 2> $OUT/${TAG}_bench_eager.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_eager.json"))
print(d["value"], d["ms_per_step"], d["e2e"], d["gpu_launches"], d.get("graph_replays"))
PY
