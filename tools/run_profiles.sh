# Round-2 evidence run (one B200): ncu launch list of the bench command (durations + DRAM bytes), full captures
# of the dominant kernels of the fp16 step and of the CTC launch(es), timings of the decode paths.
# Outputs under gpurun_out/ (copy what is to be kept into profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-twins --no-configs --no-precision-check"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 \
  --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_launches.log 2>&1
# one --set full capture of the top kernels (fp16 step): big_conv_1 forward / input gradient / weight gradient
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'wgrad_kernel' -s 57 -c 1 \
  -o gpurun_out/r02_wgrad_big1 $B > gpurun_out/r02_ncu_wgrad.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv_gemm_kernel' -s 105 -c 21 \
  -o gpurun_out/r02_conv_gemm $B > gpurun_out/r02_ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_' -c 3 -o gpurun_out/r02_ctc_bench \
  tools/selftest ctc_bench > gpurun_out/r02_ncu_ctc.log 2>&1
timeout 200 python tools/time_beam_search.py 2>&1 | tee gpurun_out/r02_beam_search_timing.log
SL_BEAM_ORDER_INDEPENDENT=1 timeout 200 python tools/time_beam_search.py 2>&1 | tee gpurun_out/r02_beam_search_timing_order_independent.log
