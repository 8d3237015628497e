set -x
cd /root/repo
mkdir -p gpurun_out
{
timeout 600 tools/selftest ctc_ | grep -E "PASS|FAIL|ms per call|bad [1-9]|finished"
for cfg in "2 4" "2 8" "2 16" "4 8" "4 16"; do
  set -- $cfg
  echo "=== SPT=$1 K=$2"
  SL_CTC_SPT=$1 SL_CTC_K=$2 timeout 120 tools/selftest ctc_bench | grep -E "ms per call|FAIL"
done
for fpb in 8 32 64; do
  echo "=== FPB=$fpb"
  SL_CTC_GRAD_FPB=$fpb timeout 120 tools/selftest ctc_bench | grep -E "ms per call"
done
for cfg in "4 8" "4 16" "8 8" "8 16"; do
  set -- $cfg
  echo "=== longform SPT=$1 K=$2"
  SL_CTC_SPT=$1 SL_CTC_K=$2 timeout 300 tools/selftest ctc_longform | grep -E "ms per call|FAIL"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_lattice_halo|ctc_grad_sorted' -c 2 -o gpurun_out/ctc_new2 tools/selftest ctc_bench | tail -3
} > gpurun_out/ctc_r3.log 2>&1
grep -v "^+" gpurun_out/ctc_r3.log
