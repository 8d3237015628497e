for ks in 0 4 5 7 8 10 12 16 21; do
  echo "== SL_WGRAD_KSPLIT=$ks"
  if [ $ks = 0 ]; then timeout 120 tools/selftest perf 64 1251 3 2>&1 | grep -E "wgrad"; else SL_WGRAD_KSPLIT=$ks timeout 120 tools/selftest perf 64 1251 3 2>&1 | grep -E "wgrad"; fi
done
