# measurement aid: phase timeline (SL_TIMELINE=1) of a conv layer's forward / input-gradient kernels under a few variants
for cfg in "SL_X=0" "SL_TAIL_KSPLIT=8" "SL_TAIL_SPLIT=0"; do
  echo "== cfg: $cfg"
  env $cfg timeout 120 tools/selftest perf 64 1251 3 2>&1 | grep -E "TFLOP|fail"
  env $cfg SL_TIMELINE=1 timeout 120 tools/selftest perf 64 1251 3 ${1:-inner_conv} 2>&1 | grep -B16 "^perf" | head -16
done
