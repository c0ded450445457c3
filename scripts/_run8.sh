m() { ncu --metrics dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pdl_main_kernel -s 3 -c 1 "$@" 2>&1 | grep -E "dram__|gpu__time|persist|cuda|max pers" | tr '\n' ' '; echo; }
echo "A launch-attr"; PDL_L2_PERSIST=1 python scripts/l2_persist_probe.py 0 heat | tail -1; PDL_L2_PERSIST=1 m python scripts/l2_persist_probe.py 0 heat
echo "B off + external stream attr"; PDL_L2_PERSIST=0 python scripts/l2_persist_probe.py 1 heat | tail -1; PDL_L2_PERSIST=0 m python scripts/l2_persist_probe.py 1 heat
echo "C internal stream attr"; PDL_L2_PERSIST=2 python scripts/l2_persist_probe.py 0 heat | tail -1; PDL_L2_PERSIST=2 m python scripts/l2_persist_probe.py 0 heat
echo "D internal stream attr, miss normal"; PDL_L2_PERSIST=3 python scripts/l2_persist_probe.py 0 heat | tail -1; PDL_L2_PERSIST=3 m python scripts/l2_persist_probe.py 0 heat
echo "E off"; PDL_L2_PERSIST=0 python scripts/l2_persist_probe.py 0 heat | tail -1; PDL_L2_PERSIST=0 m python scripts/l2_persist_probe.py 0 heat
echo "F launch-attr + external"; PDL_L2_PERSIST=1 m python scripts/l2_persist_probe.py 1 heat
echo "KS C"; PDL_L2_PERSIST=2 python scripts/l2_persist_probe.py 0 ks | tail -1; PDL_L2_PERSIST=2 m python scripts/l2_persist_probe.py 0 ks
echo "KS E"; PDL_L2_PERSIST=0 python scripts/l2_persist_probe.py 0 ks | tail -1; PDL_L2_PERSIST=0 m python scripts/l2_persist_probe.py 0 ks
