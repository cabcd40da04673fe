#!/bin/bash
# Multi-GPU pass: the 2-GPU parity test (incl. the CUDA-graph replay of the reduce-scatter exchange), then bench.py at N ranks
# with the reduce-scatter form forced (ST3R_SCATTER_FROM=2 on a 2-GPU box) with and without the graph.
set -u
TAG=${1:-r02ac}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== 2-GPU parity tests"
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short 2>&1 | tail -12 | tee $OUT/${TAG}_pytest_multi.txt
for g in 1 0; do
  echo "== bench.py --gpus $N, ST3R_TRAIN_GRAPH=$g"
  ST3R_SCATTER_FROM=2 ST3R_TRAIN_GRAPH=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$g \
      bench.py --gpus $N --steps 10 --warmup 3 --no-align --no-large --no-cpu > $OUT/${TAG}_bench_n${N}_graph$g.json 2> $OUT/${TAG}_bench_n${N}_graph$g.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n${N}_graph$g.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"], d["gpu_launches"], d.get("graph_replays"), d["roofline"]["step_breakdown_ms"], d.get("multi_gpu_parity"))
PY
  tail -3 $OUT/${TAG}_bench_n${N}_graph$g.err
done
