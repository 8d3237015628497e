#!/bin/bash
# 8-GPU run of the final code on one box: bucket split A/B, the full bench line, and one GPU of the same box.
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
for split in 1 0 1 0; do
  SL_BUCKET_SPLIT_FIRST=$split run $((29820 + RANDOM % 100)) bench.py --gpus 8 --steps 30 --warmup 5 --no-twins --no-configs --no-precision-check --no-dp-check 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('split_first=$split', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4))"
done
run 29811 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_final_n8.json 2> gpurun_out/r02_final_n8.err
python bench.py --steps 20 --warmup 5 --no-twins --no-configs --no-precision-check > gpurun_out/r02_final_n8_single.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_final_n8.json").read().strip().splitlines()[-1])
s=json.loads(open("gpurun_out/r02_final_n8_single.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ["value","ms_per_step","n_gpus","dp_check"]}, d["e2e"]["ms_per_step"])
print("single", s["ms_per_step"], s["e2e"]["ms_per_step"], "eff", d["value"]/(8*s["value"]), "e2e eff", d["e2e"]["value"]/(8*s["e2e"]["value"]))
print([(c["config"]["workload"][:30], c["ms_per_step"]) for c in d.get("configs",[])])
PY
