T=r01h
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; tail -2 gpurun_out/pytest_gpu_$T.log
python bench.py > gpurun_out/bench_${T}_heat.json 2> gpurun_out/bench_${T}_heat.err
python bench.py --workload ks --points 2097152 --no-cpu-baseline > gpurun_out/bench_${T}_ks.json 2>/dev/null
python bench.py --activation rational --no-cpu-baseline > gpurun_out/bench_${T}_heat_rational.json 2>/dev/null
python bench.py --workload burgers --points 10000 --no-cpu-baseline > gpurun_out/bench_${T}_burgers10k.json 2>/dev/null
for w in heat ks; do for on in 0 1; do PDL_L2_PERSIST=$on python scripts/probe_one.py $w Tanh 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"l2_persist=$on\", /"; done; done > gpurun_out/ab_l2_persist_$T.jsonl 2>&1
for w in heat ks; do for on in 0 1; do PDL_L2_PERSIST=$on ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pdl_main_kernel -s 2 -c 1 python scripts/ncu_target.py $w Tanh 2>&1 | grep -E "dram__|gpu__time|hit_rate" | sed "s/^/$w persist=$on /"; done; done > gpurun_out/traffic_$T.log 2>&1
PDL_L2_PERSIST=1 python bench.py --no-cpu-baseline > gpurun_out/bench_${T}_heat_l2persist.json 2>/dev/null
cat gpurun_out/traffic_$T.log; cut -c1-150 gpurun_out/ab_l2_persist_$T.jsonl
