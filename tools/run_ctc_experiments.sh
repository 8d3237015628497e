set -x
cd /root/repo
mkdir -p gpurun_out
{
timeout 900 tools/selftest ctc_ | grep -E "PASS|FAIL|ms per call|bad [1-9]|finished|rel loss|dlogits  "
echo "=== log-domain for comparison"
SL_CTC_LOG=1 timeout 300 tools/selftest ctc_bench | grep -E "ms per call|FAIL|dlogits  |rel loss"
SL_CTC_LOG=1 timeout 300 tools/selftest ctc_longform | grep -E "ms per call|FAIL|dlogits  |rel loss"
for cfg in "2 4" "2 16" "4 8" "4 16"; do
  set -- $cfg
  echo "=== bench SPT=$1 K=$2"
  SL_CTC_SPT=$1 SL_CTC_K=$2 timeout 120 tools/selftest ctc_bench | grep -E "ms per call|FAIL"
done
for cfg in "4 16 4" "4 8 4" "2 16 4" "2 8 8"; do
  set -- $cfg
  echo "=== longform SPT=$1 K=$2 cluster=$3"
  SL_CTC_SPT=$1 SL_CTC_K=$2 SL_CTC_CLUSTER=$3 timeout 300 tools/selftest ctc_longform | grep -E "ms per call|FAIL"
done
timeout 100 python tools/time_frontend.py
} > gpurun_out/ctc_r6.log 2>&1
grep -v "^+" gpurun_out/ctc_r6.log
