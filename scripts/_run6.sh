for on in 0 1; do python scripts/l2_persist_probe.py $on heat; done
for on in 0 1; do ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pdl_main_kernel -s 3 -c 2 python scripts/l2_persist_probe.py $on heat 2>&1 | grep -E "dram__|gpu__time|hit_rate|persist|cuda"; done
