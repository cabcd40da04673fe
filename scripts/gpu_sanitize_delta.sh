#!/bin/bash
# compute-sanitizer over the kernels changed since profiles/r02m_sanitize_summary.txt: surrogate-key tile sort + repair
# passes, surrogate-word correspondence merge, float4 Adam, per-direction blend batch geometry, projection (4 cameras per
# thread) and projection backward (prefetch).  Usage (GPU box, repo root): bash scripts/gpu_sanitize_delta.sh [tag]
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
export PYTHONPATH=$PWD
CS=/usr/local/cuda/bin/compute-sanitizer
run() {   # name tool pytest-args...
  local name=$1 tool=$2; shift 2
  timeout 300 $CS --tool $tool --error-exitcode 97 --print-limit 20 python -m pytest "$@" -m gpu -q -x --tb=line -p no:cacheprovider \
      > $OUT/${TAG}_sanitize_${tool}_${name}.log 2>&1
  echo "$tool $name: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' $OUT/${TAG}_sanitize_${tool}_${name}.log | tr '\n' ' ')" | tee -a $OUT/${TAG}_sanitize_summary.txt
}
: > $OUT/${TAG}_sanitize_summary.txt
run gs memcheck tests/test_gs_gpu.py -k "tile_sort_variants or fused_binning_equals_radix_chain or adam or blend_kernel_pairs_agree or train_steps or rasterization"
run gs racecheck tests/test_gs_gpu.py -k "tile_sort_variants or blend_kernel_pairs_agree or (fused_binning_equals_radix_chain and 2500)"
run merge memcheck tests/test_match_gpu.py -k "merge_corres and not 40000"
run merge racecheck tests/test_match_gpu.py -k "merge_corres and (4095 or 5000 or 16383)"
cat $OUT/${TAG}_sanitize_summary.txt
