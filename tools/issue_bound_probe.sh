# MMA issue loop: phase timeline of inner_conv and CUDA-event timings of every layer
SL_TIMELINE=1 timeout 120 tools/selftest perf 64 1251 3 inner_conv 2>&1 | grep -B16 "^perf" | grep -E "operands->tile0|tile0->tile1|tile1->tile2|issued->complete|entry->exit"
timeout 120 tools/selftest perf 64 1251 3 2>&1 | grep TFLOP
timeout 300 tools/selftest 2>&1 | grep -E "FAIL|passed|PASS" | tail -4
