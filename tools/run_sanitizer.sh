#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / TMEM / cluster pipelines: memcheck,
# racecheck and synccheck on small selftest cases of every kernel family, incl. the CTA-pair conv variant
# (SL_CTA2), the cluster-split CTC lattices, the fused CTC launch and both beam-search modes.
#   tools/run_sanitizer.sh [out_dir]
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/r02_sanitizer}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
CASES_CONV="fwd_k3_64x64 fwd_striding_k48_s2_p1 fwd_out_k1_250_33_p3 dgrad_k1_64x64 dgrad_inner_k7_250_p3 wgrad_k1_64x128 wgrad_pair_k5_128x250_p1"
CASES_CTC="ctc_small ctc_german"
summary="$OUT/summary.txt"
: > "$summary"
run() {  # name, env..., -- case
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  for tool in memcheck racecheck synccheck; do
    local log="$OUT/${name}_${tool}.log"
    env "${envs[@]}" timeout 900 $CS --tool $tool --print-limit 20 tools/selftest "$@" > "$log" 2>&1
    local errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    local fails=$(grep -c "^\[FAIL\]" "$log")
    echo "$name $tool: ${errs:-no summary line} ; selftest FAIL lines: $fails" | tee -a "$summary"
  done
}
for c in $CASES_CONV; do run "conv_$c" SL_DUMMY=0 -- "$c"; done
run conv_pair_fwd SL_CTA2=1 -- fwd_pair_odd_k3_250
run conv_pair_dgrad SL_CTA2=1 -- dgrad_pair_odd_k3_250
for c in $CASES_CTC; do run "ctc_fused_$c" SL_CTC_FUSED=1 -- "$c"; run "ctc_two_launch_$c" SL_CTC_FUSED=0 -- "$c"; done
run ctc_cluster SL_CTC_SPT=4 SL_CTC_K=16 SL_CTC_CLUSTER=2 -- ctc_german
python -m pytest tests/test_beam_search.py -q -m gpu -k "cuda_beam_search and 60-6-4" > "$OUT/beam_pytest_plain.log" 2>&1
for tool in memcheck racecheck; do
  timeout 1200 $CS --tool $tool --print-limit 20 python -m pytest tests/test_beam_search.py -q -m gpu -k "cuda_beam_search and 60-6-4" > "$OUT/beam_${tool}.log" 2>&1
  echo "beam_search $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/beam_${tool}.log" | tail -1) ; $(tail -1 "$OUT/beam_${tool}.log")" | tee -a "$summary"
done
cat "$summary"
