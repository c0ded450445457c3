#!/bin/bash
# Round-2 closing run on ONE B200 (each step under its own timeout): the full -m gpu suite, smoke(), both bench arms, then -- only
# after those exited 0 without ncu -- the launch list of the bench command, one --set full capture of pdl_main_kernel<true> and of
# pdl_fwd_tc_kernel (heat, 2^20 rows), and the DRAM-traffic metric pass bench.py's roofline.traffic is recorded from.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu > $O/r02_final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_final_smoke.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_final_bench_ref.json 2> $O/r02_final_bench_ref.err; echo "bench ref rc=$?"
timeout 400 python bench.py > $O/r02_final_bench_heat.json 2> $O/r02_final_bench_heat.err; rc=$?; echo "bench rc=$rc"
[ $rc = 0 ] || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_final_launches_bench_heat.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-rational > $O/r02_final_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pdl_main_kernel -s 2 -c 1 -f -o $O/r02_final_main_heat \
  python scripts/ncu_target.py heat Tanh 1048576 > $O/r02_final_ncu_main.log 2>&1; echo "ncu main rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pdl_fwd_tc_kernel -s 2 -c 1 -f -o $O/r02_final_tcfwd_heat \
  python scripts/tc_fwd_check.py 1048576 heat:Tanh > $O/r02_final_ncu_tcfwd.log 2>&1; echo "ncu tc fwd rc=$?"
for a in Tanh Rational; do
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:pdl_main_kernel --log-file $O/r02_final_traffic_heat_$a.log python scripts/ncu_target.py heat $a 1048576 > /dev/null 2>&1; echo "traffic $a rc=$?"
done
ls -la $O/r02_final_*
