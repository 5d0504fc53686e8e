#!/bin/bash
# compute-sanitizer memcheck + racecheck over small invocations of every hand-written kernel family
# (scan / fused scan / Levenshtein / tail of round 1, the int8-sliced tcgen05 scan + bins + resolve + lookup + walk,
# conv_tc / conv_tc3, the VQ arg-min kernels incl. the cluster kernel, the PAE kernels).  Summaries go to gpurun_out/sanitizer_*.txt; copy them to profiles/ after reading.
set -u
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SEL='test_tensor_core_scan_equals_cuda_core_reference and (128-128 or 1000-384) or test_sliced_tables_vs_float64 and 1000-384 or test_lookup_walk_equals_round1_tail or test_resolve_merges_row_shards'
SEL_OLD='test_cosine_minbycode_vs_oracle and 1000-384 or test_fused_two_block_scan_equals_separate_scans and 77-128 or test_levenshtein_minbycode_vs_oracle and 500-4 or test_rank512_stable'
SEL_VQ='test_tc_single_layers_vs_torch or test_tc3_single_layers_vs_torch or test_quantise_indices_exact or test_quantise_one_launch_kernel_equals_tiled_kernel and 30 or test_quantise_fast_path_equals_float64_kernel and 960'
# (racecheck of the golden PAE case takes ~9 minutes: the plain convolution runs 7 windows x 4 layers under the tool)
SEL_PAE='test_pose2phase_and_forward_match_golden and pae_s1 or test_shared_first_convolution_equals_per_window_convolution and 3'
for tool in memcheck racecheck; do
  for name in sliced old vq pae; do
    case $name in
      sliced) files=tests/test_sliced_gpu.py; sel="$SEL";;
      old) files=tests/test_matcher_gpu.py; sel="$SEL_OLD";;
      vq) files=tests/test_vqvae_gpu.py; sel="$SEL_VQ";;
      pae) files=tests/test_pae_gpu.py; sel="$SEL_PAE";;
    esac
    log="$OUT/sanitizer_${tool}_${name}.txt"
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 \
      python -m pytest $files -q -x -k "$sel" > "$log.full" 2>&1
    rc=$?
    { echo "# compute-sanitizer --tool $tool  python -m pytest $files -k \"$sel\"   (exit code $rc)";
      grep -E "passed|failed|error|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|Race" "$log.full" | tail -25; } > "$log"
    rm -f "$log.full.tmp"
  done
done
tail -n 4 "$OUT"/sanitizer_*.txt
