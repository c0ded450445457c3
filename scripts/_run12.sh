T=r01i
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; tail -2 gpurun_out/pytest_gpu_$T.log
python bench.py > gpurun_out/bench_${T}_heat.json 2> gpurun_out/bench_${T}_heat.err
ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pdl_main_kernel -s 2 -c 1 python scripts/ncu_target.py heat Tanh 2>&1 | grep -E "dram__|gpu__time|hit_rate" > gpurun_out/traffic_$T.log 2>&1
cat gpurun_out/traffic_$T.log; tail -1 gpurun_out/bench_${T}_heat.json | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
