for fpb in 48 56 64 72 80 96 128; do echo -n "fpb=$fpb: "; SL_CTC_GRAD_FPB=$fpb timeout 100 tools/selftest ctc_bench 2>&1 | grep -E "ms per call|FAIL" | tr '\n' ' '; echo; done
