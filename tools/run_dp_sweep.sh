#!/bin/bash
# Data-parallel checks on N GPUs of one box (default 2): equivalence with the single-GPU step (tools/dp_check.py),
# then bench.py over the communicator's CTA bound, the narrowed-grid window and torch's own communicator.
#   tools/run_dp_sweep.sh [N] [out_prefix]
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=${2:-gpurun_out/r02_dp_n$N}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
echo "=== dp_check" > $OUT.log
run 29611 tools/dp_check.py >> $OUT.log 2>&1
echo "=== bench (defaults, NCCL_DEBUG=INFO on stderr)" >> $OUT.log
NCCL_DEBUG=INFO run 29612 bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-twins > ${OUT}_default.json 2> ${OUT}_default.err
grep -E "NCCL INFO.*(Channel|nranks|comm 0x.*rank|NVLS|maxCTAs|Init COMPLETE)" ${OUT}_default.err | head -40 >> $OUT.log
port=29620
for ctas in 2 4 16 32; do
  port=$((port + 1))
  SL_COMM_MAX_CTAS=$ctas run $port bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-twins --no-dp-check --no-precision-check > ${OUT}_ctas$ctas.json 2>> $OUT.err
done
port=$((port + 1))
SL_COMM_LIMITED_LAUNCHES=0 run $port bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-twins --no-dp-check --no-precision-check > ${OUT}_nolimit.json 2>> $OUT.err
port=$((port + 1))
SL_COMM_LIMITED_LAUNCHES=4 run $port bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-twins --no-dp-check --no-precision-check > ${OUT}_limit4.json 2>> $OUT.err
port=$((port + 1))
SL_OWN_COMM=0 run $port bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-twins --no-dp-check --no-precision-check > ${OUT}_torchcomm.json 2>> $OUT.err
port=$((port + 1))
run $port bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-twins --no-dp-check --no-precision-check --pipeline-update 0 > ${OUT}_nopipe.json 2>> $OUT.err
python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-twins --no-precision-check --no-cpu-baseline > ${OUT}_single.json 2>> $OUT.err
python - "$OUT" <<'PY' >> $OUT.log
import json, glob, sys
for f in sorted(glob.glob(sys.argv[1] + "_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e)
        continue
    big = {k["kernel"]: k["ms"] for k in d["kernels"] if "big_conv_1" in k["kernel"]}
    print(f.split("/")[-1], "n", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4),
          "value", round(d["value"] / 1e6, 2), "M frames/s", d["clocks"]["sm_mhz"] if d.get("clocks") else None, big,
          d.get("dp_check"))
PY
cat $OUT.log
