#!/bin/bash
# GPU pass r02ab: one-CTA sort + unique of the correspondence merge, transposed exchange layout of the register sort,
# binning grid: parity suite, pair timing (both merge variants), bench.
set -u
TAG=${1:-r02ab}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== pair timing"
timeout 300 python scripts/nn_variants.py 2>&1 | tail -1 | tee $OUT/${TAG}_nn_variants.json
echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["step_breakdown_ms"])
print(d["match"]["value"], d["match"]["e2e"]["value"], d["match"]["roofline"]["frac"], d["reconstruct"]["seconds"], d["reconstruct"]["match_smooth_descriptors"])
PY
tail -3 $OUT/${TAG}_bench.err
echo "== launch list of a matched pair"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_match_pair.csv \
    python scripts/prof_step.py match > $OUT/ncu_match.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/${TAG}_launches_match_pair.csv")))
hdr = None; agg = collections.Counter(); cnt = collections.Counter()
for r in rows:
    if len(r) > 5 and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d["Metric Value"].replace(",", ""))
        except ValueError: continue
        k = d["Kernel Name"][:48]; agg[k] += v; cnt[k] += 1
for k, v in agg.most_common(14): print(f"{v/1e3:10.1f} us {cnt[k]:5d} x  {k}")
PY
