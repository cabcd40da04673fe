#!/bin/bash
# N-GPU bench pass (gpurun --gpus N): NVLS exchange vs the peer-load exchange, then the full bench line (sharded
# reconstruct + the BASELINE.json config of this GPU count).  Usage: bash scripts/gpu_multi8.sh <tag> <N>
set -u
TAG=${1:-r02}
N=${2:-8}
OUT=gpurun_out
export PYTHONPATH=$PWD
nvidia-smi -L | wc -l
run() {  # name extra-env...
  local name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 $EXTRA > $OUT/${TAG}_bench_n${N}_$name.json 2> $OUT/${TAG}_bench_n${N}_$name.err
  tail -2 $OUT/${TAG}_bench_n${N}_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", d["config"]["parallelism"]); print(d["value"] / 1e6, "M G/s", d["ms_per_step"], d["roofline"]["step_breakdown_ms"])
    print(d["multi_gpu_parity"]); print(d.get("reconstruct")); print(d.get("baseline_config_at_this_gpu_count"))
except Exception as e:
    print("no line:", e)
PY
}
EXTRA="--no-align --no-large" run p2p ST3R_NVLS=0
EXTRA="" run nvls ST3R_NVLS=1
