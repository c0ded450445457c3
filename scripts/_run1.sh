python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01e.log 2>&1; tail -3 gpurun_out/pytest_gpu_r01e.log
for w in "heat Rational" "burgers Rational" "ks Rational" "rd2d Rational"; do set -- $w; for wg in 0 1; do PDL_WGRAD_MMA=$wg python scripts/probe_one.py $1 $2 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"wgrad_mma $wg\", /"; done; done > gpurun_out/ab_rational_wgrad.log 2>&1
python bench.py > gpurun_out/bench_r01e_heat.json 2> gpurun_out/bench_r01e_heat.err; tail -1 gpurun_out/bench_r01e_heat.json | cut -c1-400
python scripts/gpu_errors.py > gpurun_out/gpu_errors_r01e.log 2>&1; tail -1 gpurun_out/gpu_errors_r01e.log
