#!/bin/bash
# CTC fused-launch sweep on one B200: correctness of every selftest CTC case, then timings of the bench /
# long-form shapes against the two-launch path and over the gradient-CTA count and frames per item.
cd "$(dirname "$0")/.."
echo "=== fused (default), every ctc case"
timeout 600 tools/selftest ctc_ 2>&1 | grep -E 'PASS|FAIL|ms per call|finished'
echo "=== two launches (SL_CTC_FUSED=0)"
SL_CTC_FUSED=0 timeout 600 tools/selftest ctc_bench 2>&1 | grep -E 'PASS|FAIL|ms per call'
SL_CTC_FUSED=0 timeout 600 tools/selftest ctc_longform 2>&1 | grep -E 'PASS|FAIL|ms per call'
for ctas in 74 148 296 444 592; do
  for fpi in 8 16 32; do
    echo "=== ctc_bench SL_CTC_GRAD_CTAS=$ctas SL_CTC_GRAD_FPI=$fpi"
    SL_CTC_GRAD_CTAS=$ctas SL_CTC_GRAD_FPI=$fpi timeout 300 tools/selftest ctc_bench 2>&1 | grep -E 'FAIL|gradient: '
  done
done
for ctas in 32 64 128; do
  for fpi in 16 64; do
    echo "=== ctc_longform SL_CTC_GRAD_CTAS=$ctas SL_CTC_GRAD_FPI=$fpi"
    SL_CTC_GRAD_CTAS=$ctas SL_CTC_GRAD_FPI=$fpi timeout 300 tools/selftest ctc_longform 2>&1 | grep -E 'FAIL|gradient: '
  done
done
