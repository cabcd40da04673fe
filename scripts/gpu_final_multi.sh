#!/bin/bash
# Multi-GPU confirmation of a build (gpurun --gpus N): the 2-GPU parity tests, then the full bench line at N GPUs.
# Usage: bash scripts/gpu_final_multi.sh <tag> <N>
set -u
TAG=${1:-r02}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --tb=short 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
tail -3 $OUT/${TAG}_bench_n$N.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print(d["config"]["parallelism"]); print(d["value"] / 1e6, "M G/s", d["ms_per_step"], d["roofline"]["step_breakdown_ms"])
print(d["multi_gpu_parity"]); print(d.get("reconstruct")); print(d.get("baseline_config_at_this_gpu_count"))
PY
