# CTC tuning sweep on one B200 (gpurun -- 'bash tools/run_ctc_experiments.sh'): parity + CUDA-event timings
# of every selftest CTC case with the default configuration, then the tuning variables of
# speechless_b200/csrc/ctc.cu at the bench shape (64 x 626 frames, S = 301) and at the long-form shape
# (16 x 3751 frames, S = 1801).  SL_CTC_SPT = states per lane, SL_CTC_K = steps per barrier,
# SL_CTC_CLUSTER = CTAs per lattice, SL_CTC_GRAD_FPB = frames per gradient block, SL_CTC_LEGACY = 1 / 2:
# first-generation kernels.  Output: gpurun_out/ctc_sweep.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
timeout 900 tools/selftest ctc_ | grep -E "PASS|FAIL|ms per call|bad [1-9]|finished"
for cfg in "2 4" "2 8" "2 16" "4 8" "4 16"; do
  set -- $cfg
  echo "=== bench shape SPT=$1 K=$2"
  SL_CTC_SPT=$1 SL_CTC_K=$2 timeout 120 tools/selftest ctc_bench | grep -E "ms per call|FAIL"
done
for fpb in 8 16 32 64; do
  echo "=== bench shape gradient frames per block = $fpb"
  SL_CTC_GRAD_FPB=$fpb timeout 120 tools/selftest ctc_bench | grep -E "ms per call|FAIL"
done
for legacy in 1 2; do
  echo "=== bench shape SL_CTC_LEGACY=$legacy"
  SL_CTC_LEGACY=$legacy timeout 120 tools/selftest ctc_bench | grep -E "ms per call|FAIL"
done
for cfg in "4 8 1" "4 8 2" "4 8 4" "4 16 4" "4 32 4" "2 16 4" "8 32 4"; do
  set -- $cfg
  echo "=== long-form shape SPT=$1 K=$2 cluster=$3"
  SL_CTC_SPT=$1 SL_CTC_K=$2 SL_CTC_CLUSTER=$3 timeout 300 tools/selftest ctc_longform | grep -E "ms per call|FAIL"
done
} > gpurun_out/ctc_sweep.log 2>&1
cat gpurun_out/ctc_sweep.log
