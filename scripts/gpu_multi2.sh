set -u
export PYTHONPATH=$PWD
OUT=gpurun_out
ST3R_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_experimental_gpu.py -m gpu -q --tb=short -s 2>&1 | grep -E "NVLS|passed|failed|Error|assert" | tail -12 | tee $OUT/r02r_pytest_multi.txt
for sf in 2 4; do
ST3R_SCATTER_FROM=$sf timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-align > $OUT/r02r_bench_n2_sf$sf.json 2> $OUT/r02r_bench_n2_sf$sf.err
tail -3 $OUT/r02r_bench_n2_sf$sf.err
python - <<PY
import json
d = json.loads(open("$OUT/r02r_bench_n2_sf$sf.json").read().strip().splitlines()[-1])
print("scatter_from=$sf", d["config"]["parallelism"]); print(d["ms_per_step"], d["roofline"]["step_breakdown_ms"]); print(d["multi_gpu_parity"])
PY
done
