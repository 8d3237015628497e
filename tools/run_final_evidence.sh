python bench.py > gpurun_out/r02_bench_v5.json 2> gpurun_out/r02_bench_v5.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_v5_reference.json 2>/dev/null
CS=/usr/local/cuda/bin/compute-sanitizer
for c in fwd_k3_64x64 dgrad_inner_k7_250_p3 wgrad_pair_k5_128x250_p1 fwd_striding_k48_s2_p1; do
  for tool in racecheck memcheck; do
    echo "$c $tool: $(timeout 300 $CS --tool $tool --print-limit 5 tools/selftest $c 2>&1 | grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|FAIL' | tr '\n' ' ')"
  done
done
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-twins --no-configs --no-precision-check"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_final2.csv $B > /dev/null 2>&1
SL_TIMELINE=1 timeout 120 tools/selftest perf 64 1251 3 inner_conv 2>&1 | grep -B16 "^perf" | head -16
SL_TIMELINE=1 timeout 120 tools/selftest perf 64 1251 3 big_conv_1 2>&1 | grep -B16 "^perf" | head -16
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_v5.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'], d['roofline']['all_conv'], d['precision']['meets_tolerance'], d['precision']['logits_rel_err'], [t['ms_per_step'] for t in d['twins']], [c['ms_per_step'] for c in d['configs']], d['clocks'])"
