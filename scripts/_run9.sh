T=r01g
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; tail -2 gpurun_out/pytest_gpu_$T.log
for w in "heat Tanh" "ks Tanh" "rd2d Rational" "burgers Tanh"; do set -- $w; for on in 1 0; do PDL_L2_PERSIST=$on python scripts/probe_one.py $1 $2 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"l2_persist=$on\", /"; done; done > gpurun_out/ab_l2_persist_$T.jsonl 2>&1
for w in "heat Tanh" "ks Tanh" "heat Rational"; do set -- $w; ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pdl_main_kernel -s 2 -c 1 python scripts/ncu_target.py $1 $2 2>&1 | grep -E "dram__|gpu__time|hit_rate" | sed "s/^/$1 $2 /"; done > gpurun_out/traffic_$T.log 2>&1
python bench.py > gpurun_out/bench_${T}_heat.json 2> gpurun_out/bench_${T}_heat.err
python bench.py --workload ks --points 2097152 --no-cpu-baseline > gpurun_out/bench_${T}_ks.json 2>/dev/null
python scripts/epoch_rate.py > gpurun_out/epoch_rate_$T.log 2>&1
ncu --metrics dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pdl_main_kernel -s 40 -c 3 python scripts/epoch_rate.py 2>&1 | grep -E "dram__|gpu__time" > gpurun_out/traffic_graph_$T.log
cat gpurun_out/traffic_$T.log gpurun_out/traffic_graph_$T.log
