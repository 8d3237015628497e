#!/bin/bash
# 8-GPU run (one box): the default bench line (with dp_check and the other configurations), then the
# communicator's CTA bound, the narrowed-grid window off, and torch's own communicator for comparison.
cd "$(dirname "$0")/.."
N=${1:-8}
OUT=gpurun_out/r02_dp_n$N
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
NCCL_DEBUG=INFO run 29712 bench.py --gpus $N --steps 30 --warmup 5 --no-twins > ${OUT}_default.json 2> ${OUT}_default.err
grep -E "NCCL INFO.*(nranks|NVLS|Init COMPLETE|ncclCommInitRank)" ${OUT}_default.err | head -30 > $OUT.log
S="--steps 30 --warmup 5 --no-configs --no-twins --no-dp-check --no-precision-check"
SL_COMM_MAX_CTAS=16 run 29713 bench.py --gpus $N $S > ${OUT}_ctas16.json 2>> $OUT.err
SL_COMM_MAX_CTAS=4 run 29714 bench.py --gpus $N $S > ${OUT}_ctas4.json 2>> $OUT.err
SL_COMM_MAX_CTAS=24 SL_COMM_LIMITED_LAUNCHES=3 run 29715 bench.py --gpus $N $S > ${OUT}_ctas24.json 2>> $OUT.err
SL_COMM_LIMITED_LAUNCHES=0 run 29716 bench.py --gpus $N $S > ${OUT}_nolimit.json 2>> $OUT.err
SL_OWN_COMM=0 run 29717 bench.py --gpus $N $S > ${OUT}_torchcomm.json 2>> $OUT.err
python bench.py --gpus 1 --steps 30 --warmup 5 --no-configs --no-twins --no-precision-check --no-cpu-baseline > ${OUT}_single.json 2>> $OUT.err
python - "$OUT" <<'PY' >> $OUT.log
import json, glob, sys
for f in sorted(glob.glob(sys.argv[1] + "_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e)
        continue
    print(f.split("/")[-1], "n", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 4),
          "value", round(d["value"] / 1e6, 2), "M frames/s", d["clocks"]["sm_mhz"] if d.get("clocks") else None, d.get("dp_check"))
    for c in d.get("configs") or []:
        print("    config", c["config"]["workload"][:40], c["config"]["global_batch"], c["ms_per_step"], round(c["value"] / 1e6, 2), c["ctc"])
PY
cat $OUT.log
