#!/bin/bash
# compute-sanitizer pass over the small cases of the -m gpu suite: memcheck on every path,
# racecheck on the shared-memory heavy kernels (blend forward / backward, tile binning, ALIGN), synccheck on the
# tcgen05 matcher.  Usage (GPU box, repo root): bash scripts/gpu_sanitize.sh [tag]    (~10 GPU-minutes)
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
CS=/usr/local/cuda/bin/compute-sanitizer
SMALL_GS='rasterization or blend_kernel_pairs_agree or visit_list or loss_forward or fused_adam or train_steps or train_plan or fused_binning_equals_radix_chain'
run() {   # name tool pytest-args...
  local name=$1 tool=$2; shift 2
  echo "== $tool: $name"
  timeout 1500 $CS --tool $tool --error-exitcode 97 --print-limit 20 python -m pytest "$@" -m gpu -q -x --tb=line -p no:cacheprovider \
      > $OUT/${TAG}_sanitize_${tool}_${name}.log 2>&1
  echo "rc=$? $(grep -E 'ERROR SUMMARY|passed|failed' $OUT/${TAG}_sanitize_${tool}_${name}.log | tr '\n' ' ')" | tee -a $OUT/${TAG}_sanitize_summary.txt
}
: > $OUT/${TAG}_sanitize_summary.txt
run gs memcheck tests/test_gs_gpu.py -k "$SMALL_GS"
run gs racecheck tests/test_gs_gpu.py -k "rasterization or blend_kernel_pairs_agree or fused_binning_equals_radix_chain"
run match memcheck tests/test_match_gpu.py -k "golden or vs_oracle"
run match synccheck tests/test_match_gpu.py -k "nn_argmax_golden or extract_correspondences_golden"
run align memcheck tests/test_align_gpu.py
run align racecheck tests/test_align_gpu.py -k "kernel_loss_and_gradients or canonical_view"
run mcmc memcheck tests/test_mcmc_gpu.py
cat $OUT/${TAG}_sanitize_summary.txt
