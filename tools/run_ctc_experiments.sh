set -x
cd /root/repo
mkdir -p gpurun_out
{
timeout 900 tools/selftest ctc_ | grep -E "PASS|FAIL|ms per call|bad [1-9]|finished"
for cfg in "4 32 4" "8 32 4" "8 16 4" "4 16 2"; do
  set -- $cfg
  echo "=== longform SPT=$1 K=$2 cluster=$3"
  SL_CTC_SPT=$1 SL_CTC_K=$2 SL_CTC_CLUSTER=$3 timeout 300 tools/selftest ctc_longform | grep -E "ms per call|FAIL"
done
} > gpurun_out/ctc_r5.log 2>&1
grep -v "^+" gpurun_out/ctc_r5.log
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py 2>&1 | tail -1 > gpurun_out/bench_full.json; cut -c1-400 gpurun_out/bench_full.json
python bench.py --workload longform --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_longform.json; cut -c1-300 gpurun_out/bench_longform.json
python bench.py --workload small --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_small.json; cut -c1-300 gpurun_out/bench_small.json
