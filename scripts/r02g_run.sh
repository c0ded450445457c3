cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > $O/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02g_pytest_gpu.log
for f in "" "-DPDL_NO_POINT_PREFETCH" ""; do echo "== flags: $f"; PDL_NVCC_FLAGS="$f" PDL_PROBE_FLUSH=1 timeout 400 python scripts/perf_probe.py 2>&1 | grep "^{" | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if 'ms' in d: print('  %-8s %-8s B=%-8d bwd=%d  %.4f ms (min %.4f)' % (d['workload'], d['act'], d['B'], d['bwd'], d['ms'], d['min_ms']))
"; done > $O/r02g_ab_prefetch.txt 2>&1; cat $O/r02g_ab_prefetch.txt
timeout 400 python bench.py > $O/r02g_bench_heat.json 2> $O/r02g_bench_heat.err; echo "bench rc=$?"
for a in Tanh Rational; do
  timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
    -k regex:pdl_main_kernel --log-file $O/r02g_traffic_heat_$a.log python scripts/ncu_target.py heat $a 1048576 > /dev/null 2>&1; echo "traffic $a rc=$?"; tail -5 $O/r02g_traffic_heat_$a.log
done
