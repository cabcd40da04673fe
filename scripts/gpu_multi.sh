#!/bin/bash
# Multi-GPU pass on one box (gpurun --gpus N): the 2-GPU parity tests, then bench.py at N ranks (its own parity checks run
# before anything is timed).  Usage: bash scripts/gpu_multi.sh <tag> <N>
set -u
TAG=${1:-r02}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
nvidia-smi -L | tee $OUT/${TAG}_gpus.txt
echo "== 2-GPU parity tests"
ST3R_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_experimental_gpu.py -m gpu -q --tb=short 2>&1 | tail -12 | tee $OUT/${TAG}_pytest_multi.txt
echo "== bench.py --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
tail -c 3000 $OUT/${TAG}_bench_n$N.json; tail -5 $OUT/${TAG}_bench_n$N.err
