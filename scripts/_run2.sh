run() { PDL_WGRAD_MMA=$1 PDL_NVCC_FLAGS="$2" python scripts/probe_one.py $3 $4 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"wgrad_mma=$1 $2\", /"; }
for w in "heat Rational" "burgers Rational" "ks Rational" "rd2d Rational"; do set -- $w; run "" "-DPDL_RCP_IEEE" $1 $2; run "" "" $1 $2; run 1 "" $1 $2; done > gpurun_out/ab_rational_rcp.log 2>&1
python scripts/gpu_errors.py > gpurun_out/gpu_errors_r01e.log 2>&1; tail -1 gpurun_out/gpu_errors_r01e.log
python scripts/ncu_target.py heat Tanh > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pdl_main_kernel -s 2 -c 1 -f -o gpurun_out/prof_heat_r01e python scripts/ncu_target.py heat Tanh > gpurun_out/ncu_r01e.log 2>&1; tail -2 gpurun_out/ncu_r01e.log
python bench.py --workload ks --points 2097152 --no-cpu-baseline > gpurun_out/bench_r01e_ks.json 2>gpurun_out/bench_r01e_ks.err; tail -1 gpurun_out/bench_r01e_ks.json | cut -c1-200
