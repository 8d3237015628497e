# Round-1 evidence run (one B200): GPU tests, bench lines, ncu launch list of the bench command,
# full captures of the CTC kernels.  Outputs under gpurun_out/ (copy what is to be kept into profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_full.json; cut -c1-200 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json; cut -c1-200 gpurun_out/bench_reference.json
timeout 600 python bench.py --workload long --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_long.json; cut -c1-200 gpurun_out/bench_long.json
timeout 600 python bench.py --workload small --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_small.json; cut -c1-200 gpurun_out/bench_small.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 200 python tools/time_beam_search.py 2>&1 | tee gpurun_out/beam_search_timing.log
timeout 100 python tools/time_frontend.py 2>&1 | tee gpurun_out/frontend_timing.log
if [ "$1" = "ncu-ctc" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_lattice_halo|ctc_grad_sorted' -c 2 -o gpurun_out/ctc_bench tools/selftest ctc_bench | tail -3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_lattice_halo|ctc_grad_sorted' -c 2 -o gpurun_out/ctc_longform tools/selftest ctc_longform | tail -3
fi
