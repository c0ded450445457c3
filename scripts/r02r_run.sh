#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out
timeout 800 python -m pytest tests -x -q -m gpu > $O/r02r_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -v Warn $O/r02r_pytest_gpu.log | tail -2
PDL_PROBE_FLUSH=1 timeout 300 python scripts/perf_probe.py 2>&1 | grep "^{" > $O/r02r_perf_probe.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/r02r_perf_probe.jsonl"):
    d = json.loads(l)
    if "ms" in d: print("  %-8s %-8s B=%-8d bwd=%d  %.4f ms  %.3e pts/s  %.1f TFLOP/s" % (d["workload"], d["act"], d["B"], d["bwd"], d["ms"], d["pts_per_s"], d["tflops"]))
    else: print(d)
PY
rm -f $O/r02r_bench_n1.jsonl
for w in kdv burgers; do timeout 250 python bench.py --workload $w --scaling strong --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" >> $O/r02r_bench_n1.jsonl; done
python - <<'PY'
import json
for l in open("gpurun_out/r02r_bench_n1.jsonl"):
    d = json.loads(l); print("%-50s %.4e pts/s %.3f ms/step e2e %.4e kernel %.3f | rational %.4e" % (d["config"]["workload"][:50], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["rational"]["value"]))
PY
