# CTC loss + gradient at the long-form shape (16 x 60 s): frames per block / staging of the gradient kernel
for st in 1 0; do for fpb in 16 32 64 128; do
  echo -n "staged=$st fpb=$fpb: "; SL_CTC_GRAD_STAGED=$st SL_CTC_GRAD_FPB=$fpb timeout 100 tools/selftest ctc_longform 2>&1 | grep -E "ms per call|FAIL" | head -2 | tr '\n' ' '; echo
done; done
echo -n "default: "; timeout 100 tools/selftest ctc_longform 2>&1 | grep -E "ms per call|FAIL|PASS" | tr '\n' ' '; echo
