# measurement aid: phase timeline (SL_TIMELINE=1) and timings of a conv layer's forward / input-gradient kernels under a few variants
for cfg in "SL_X=0" "SL_DBG_MODE=1" "SL_DBG_MODE=2" "SL_DBG_MODE=3" "SL_HALO=1"; do
  echo "== cfg: $cfg"
  env $cfg SL_TIMELINE=1 timeout 120 tools/selftest perf 64 1251 3 ${1:-inner_conv} 2>&1 | grep -B16 "^perf" | grep -E "operands->tile0|tile0->tile1|issued->complete|entry->exit" | head -4
  env $cfg timeout 120 tools/selftest perf 64 1251 3 2>&1 | grep -E "TFLOP" | grep -E "fwd|dgrad"
done
