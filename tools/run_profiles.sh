# Round-1 evidence run (one B200): GPU tests, bench lines, ncu launch list of the bench command,
# full captures of the CTC kernels.  Outputs under gpurun_out/.
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 200 python tools/time_beam_search.py 2>&1 | tee gpurun_out/beam_search_timing.log
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_v10.json; cut -c1-300 gpurun_out/bench_v10.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_v10_reference.json; cut -c1-300 gpurun_out/bench_v10_reference.json
timeout 600 python bench.py --workload long --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_v10_long.json; cut -c1-300 gpurun_out/bench_v10_long.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_v4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_v4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_lattice_halo|ctc_grad_sorted' -c 2 -o gpurun_out/ctc_v4_bench tools/selftest ctc_bench | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_lattice_halo|ctc_grad_sorted' -c 2 -o gpurun_out/ctc_v4_longform tools/selftest ctc_longform | tail -3
timeout 300 tools/selftest ctc_ 2>&1 | grep -E "PASS|FAIL|ms per call|finished" > gpurun_out/selftest_ctc_v4.log; tail -3 gpurun_out/selftest_ctc_v4.log
