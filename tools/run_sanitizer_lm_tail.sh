#!/bin/bash
# compute-sanitizer over the kernels added late in round 2: the language-model beam search
# (sl_ctc_beam_search_decode_lm) and the opt-in tail K split of the conv GEMM (SL_TAIL_KSPLIT), driven
# through their pytest cases.   tools/run_sanitizer_lm_tail.sh [out_dir]
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/r02_sanitizer}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
summary="$OUT/summary_lm_tail.txt"
: > "$summary"
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 python -m pytest tests/test_beam_search.py -q -m gpu \
      -k "language_model_matches and (abc-3-4 or ab-2-1)" > "$OUT/beam_lm_${tool}.log" 2>&1
  echo "beam_search_lm $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/beam_lm_${tool}.log" | tail -1) ; $(tail -1 "$OUT/beam_lm_${tool}.log")" | tee -a "$summary"
  timeout 900 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_bench_variants.py -q -m gpu \
      -k "test_forward_tail_split_tiles and k-split and fp16 and 250-250-7" > "$OUT/conv_tail_ksplit_${tool}.log" 2>&1
  echo "conv_tail_ksplit $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$OUT/conv_tail_ksplit_${tool}.log" | tail -1) ; $(tail -1 "$OUT/conv_tail_ksplit_${tool}.log")" | tee -a "$summary"
done
cat "$summary"
