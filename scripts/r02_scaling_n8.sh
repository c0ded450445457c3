#!/bin/bash
# Trimmed re-run of the N = 8 points of scripts/r02_scaling_run.sh with the final kernels (fp16 forward cross terms).
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="--steps 10 --warmup 3 --no-cpu-baseline"
for args in "--workload ks --scaling strong" "--workload rd2d --scaling strong" "--workload kdv --scaling strong" "--workload heat"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 $B $args \
    2>> $O/r02p_scale_n8.err | grep "^{" >> $O/r02p_scale_n8.jsonl; echo "n=8 $args rc=${PIPESTATUS[0]}"
done
python - <<'PY'
import json
for l in open("gpurun_out/r02p_scale_n8.jsonl"):
    r = json.loads(l); rat = r.get("rational") or {}
    print("%-52s N=%d %-6s rows_total=%-9d %.4e pts/s  %.2f ms/step  e2e %.4e | rational %.4e" % (
        r["config"]["workload"][:52], r["n_gpus"], r["scaling"], r["config"]["rows_total"], r["value"], r["ms_per_step"], r["e2e"]["value"], rat.get("value", float("nan"))))
PY
