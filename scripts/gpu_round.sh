#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, ncu launch lists and full captures of the top kernels (the 20th
# training step, i.e. the regime bench.py times), and the traffic / issue-slot JSON bench.py embeds - all from the same
# build.  Usage (from the repo root on the GPU box): bash scripts/gpu_round.sh [tag]       (~8 GPU-minutes)
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== ncu full captures (step 20 of the headline workload; the matcher at M = 4096 / 262144)"
for k in raster_bwd_pool_kernel raster_fwd_pool_kernel tile_sort_kernel ssim_l1_fwd_kernel; do
  ST3R_PROF_STEPS=21 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 19 -c 1 -f -o $OUT/${TAG}_prof_$k \
      python scripts/prof_step.py step > $OUT/ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_tc_kernel --launch-skip 1 -c 1 -f -o $OUT/${TAG}_prof_nn_tc_M4096 \
    python scripts/prof_nn.py tcgen05 4096 > $OUT/ncu_nn_tc.log 2>&1
python scripts/ncu_to_json.py $OUT/${TAG}_traffic.json st3r_gs_raster_bwd=$OUT/${TAG}_prof_raster_bwd_pool_kernel.ncu-rep \
    st3r_gs_raster_fwd=$OUT/${TAG}_prof_raster_fwd_pool_kernel.ncu-rep tile_sort_kernel=$OUT/${TAG}_prof_tile_sort_kernel.ncu-rep \
    st3r_gs_loss_fwd=$OUT/${TAG}_prof_ssim_l1_fwd_kernel.ncu-rep nn_tc_kernel=$OUT/${TAG}_prof_nn_tc_M4096.ncu-rep > /dev/null
python scripts/ncu_summary.py $OUT/${TAG}_prof_*.ncu-rep > $OUT/${TAG}_ncu_summary.txt 2>&1
mkdir -p profiles && cp $OUT/${TAG}_traffic.json profiles/traffic.json     # what bench.py reads (same build, same pass)
echo "== bench"
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 6000 $OUT/${TAG}_bench.json
tail -5 $OUT/${TAG}_bench.err
echo "== ncu launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_train_step.csv \
    python scripts/prof_step.py step > $OUT/ncu_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_match_pair.csv \
    python scripts/prof_step.py match > $OUT/ncu_match.log 2>&1
echo "== MATCH microbenchmark (BASELINE configs[4]: tcgen05 vs SIMT vs cuBLAS GEMM + max, measured peaks)"
timeout 600 python scripts/bench_match.py > $OUT/${TAG}_match_micro.json 2> $OUT/${TAG}_match_micro.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_match_micro.json"))
print(d["peaks_tflops"], {k: round(v["tcgen05"]["tflops"], 1) for k, v in d["nn_argmax"].items()}, d["all_pairs_tcgen05"])
PY
echo "== SASS evidence (tcgen05 / TMA / TMEM mnemonics of the shipped library)"
cuobjdump -sass starst3r_b200/libstarst3r_b200.so 2>/dev/null | grep -oE "UTCHMMA[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|UTCATOMSWS[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|ATOMS\.OR|REDG\.E\.ADD\.F32x4[A-Z0-9_.]*" \
    | sort | uniq -c | sort -rn > $OUT/${TAG}_sass_mnemonics.txt
cat $OUT/${TAG}_sass_mnemonics.txt
ls -la $OUT | tail -30
