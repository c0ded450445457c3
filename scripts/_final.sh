T=r01f
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; tail -2 gpurun_out/pytest_gpu_$T.log
python bench.py > gpurun_out/bench_${T}_heat.json 2> gpurun_out/bench_${T}_heat.err
python bench.py --workload ks --points 2097152 --no-cpu-baseline > gpurun_out/bench_${T}_ks.json 2>/dev/null
python bench.py --activation rational --no-cpu-baseline > gpurun_out/bench_${T}_heat_rational.json 2>/dev/null
python bench.py --workload burgers --points 10000 --no-cpu-baseline > gpurun_out/bench_${T}_burgers10k.json 2>/dev/null
python bench.py --workload kdv --points 2097152 --no-cpu-baseline > gpurun_out/bench_${T}_kdv.json 2>/dev/null
python scripts/gpu_errors.py > gpurun_out/gpu_errors_$T.log 2>&1; tail -1 gpurun_out/gpu_errors_$T.log
bash scripts/ab_engine.sh > gpurun_out/ab_engines_$T.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}_bench_heat.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pdl_main_kernel -s 2 -c 1 -f -o gpurun_out/prof_heat_$T python scripts/ncu_target.py heat Tanh > gpurun_out/ncu_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pdl_main_kernel -s 2 -c 1 -f -o gpurun_out/prof_ks_$T python scripts/ncu_target.py ks Tanh >> gpurun_out/ncu_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pdl_main_kernel -s 2 -c 1 -f -o gpurun_out/prof_heat_rational_$T python scripts/ncu_target.py heat Rational >> gpurun_out/ncu_$T.log 2>&1
tail -2 gpurun_out/ncu_$T.log
