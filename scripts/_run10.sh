run() { PDL_L2_PERSIST=$1 PDL_L2_PART=$2 PDL_L2_RATIO=$3 python scripts/probe_one.py $4 Tanh 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"persist=$1 part=$2 ratio=$3\", /"; }
for w in heat ks; do run 0 "" "" $w; run 1 "" "" $w; run 1 s "" $w; run 1 w "" $w; run 1 "" 0.8 $w; run 1 "" 0.5 $w; done
run 0 "" "" heat; run 1 "" "" heat
