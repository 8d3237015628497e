#!/bin/bash
# CTC fused-launch sweep on one B200: correctness of every selftest CTC case, then timings of the bench /
# long-form shapes against the two-launch path over the number of gradient CTAs, and the measurement aids.
cd "$(dirname "$0")/.."
echo "=== default all cases"; timeout 600 tools/selftest ctc_ 2>&1 | grep -E 'PASS|FAIL|gradient: |finished'
echo "=== two launches"; SL_CTC_FUSED=0 timeout 300 tools/selftest ctc_bench 2>&1 | grep -E 'ms per call'
SL_CTC_FUSED=0 timeout 300 tools/selftest ctc_longform 2>&1 | grep -E 'ms per call'
for ctas in 64 128 148 212 296 464; do echo "=== ctc_bench CTAS=$ctas"; SL_CTC_GRAD_CTAS=$ctas timeout 300 tools/selftest ctc_bench 2>&1 | grep -E 'FAIL|gradient: '; done
for d in 1 2 3; do echo "=== ctc_bench DBG=$d"; SL_CTC_DBG=$d timeout 300 tools/selftest ctc_bench 2>&1 | grep -E 'gradient: '; done
for ctas in 64 128 168; do echo "=== ctc_longform CTAS=$ctas"; SL_CTC_GRAD_CTAS=$ctas timeout 300 tools/selftest ctc_longform 2>&1 | grep -E 'FAIL|gradient: '; done
