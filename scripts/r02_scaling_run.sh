#!/bin/bash
# Round-2 multi-GPU evidence on ONE 8-GPU box (gpurun --gpus 8): sharded-vs-single parity at world 2/4/8 (eager step and captured
# epochs) and the strong-scaling points of BASELINE.json configs 3-5 with the final kernels ("next round" item 1).
# Every command runs under its own timeout.  Output: gpurun_out/r02_scale_*.{log,jsonl}
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
B="--steps 10 --warmup 3 --no-cpu-baseline"
run() {  # run <n_gpus> <visible devices> <port> <tag> <bench args...>
  local n=$1 vis=$2 port=$3 tag=$4; shift 4
  if [ "$n" = 1 ]; then CUDA_VISIBLE_DEVICES=$vis timeout 300 python bench.py --gpus 1 $B "$@" >> $O/r02_scale_$tag.jsonl 2>> $O/r02_scale_$tag.err
  else CUDA_VISIBLE_DEVICES=$vis timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
       bench.py --gpus $n $B "$@" >> $O/r02_scale_$tag.jsonl 2>> $O/r02_scale_$tag.err; fi
  echo "[$tag] n=$n $* rc=$?"
}
# ---- phase A: everything that needs all 8 GPUs ----
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -s -k "8" > $O/r02_scale_pytest_world8.log 2>&1; echo "pytest world8 rc=$?"
run 8 0,1,2,3,4,5,6,7 29501 n8 --workload ks --scaling strong
run 8 0,1,2,3,4,5,6,7 29502 n8 --workload rd2d --scaling strong
run 8 0,1,2,3,4,5,6,7 29503 n8 --workload heat
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29504 tests/dist_graph_worker.py --time \
  > $O/r02_scale_graph_dp8.log 2>&1; echo "graph dp8 rc=$?"
# ---- phase B: N = 4, 2, 1 side by side on disjoint GPUs ----
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -s -k "4" > $O/r02_scale_pytest_world4.log 2>&1; echo "pytest world4 rc=$?"
  run 4 0,1,2,3 29511 n4 --workload kdv --scaling strong
  run 4 0,1,2,3 29512 n4 --workload rd2d --scaling strong
  run 4 0,1,2,3 29513 n4 --workload heat ) &
( CUDA_VISIBLE_DEVICES=4,5 timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -s -k "2" > $O/r02_scale_pytest_world2.log 2>&1; echo "pytest world2 rc=$?"
  run 2 4,5 29521 n2 --workload kdv --scaling strong
  run 2 4,5 29522 n2 --workload rd2d --scaling strong
  run 2 4,5 29523 n2 --workload heat ) &
( run 1 6 0 n1 --workload ks --scaling strong
  run 1 6 0 n1 --workload heat ) &
( run 1 7 0 n1b --workload rd2d --scaling strong
  run 1 7 0 n1b --workload kdv --scaling strong ) &
wait
cat $O/r02_scale_n*.jsonl > $O/r02_scale_all.jsonl
python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/r02_scale_all.jsonl") if l.startswith("{")]
for r in sorted(rows, key=lambda r: (r["config"]["workload"], r["n_gpus"])):
    rat = r.get("rational") or {}
    print("%-52s N=%d %-6s rows_total=%-9d %.3e pts/s  %.2f ms/step  e2e %.3e | rational %.3e" % (
        r["config"]["workload"][:52], r["n_gpus"], r["scaling"], r["config"]["rows_total"], r["value"], r["ms_per_step"],
        r["e2e"]["value"], rat.get("value", float("nan"))))
PY
