run() { PDL_ADJ_FUSED=$1 PDL_NVCC_FLAGS="$2" python scripts/probe_one.py $3 $4 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"fused=$1 $2\", /"; }
for w in "heat Tanh" "kdv Tanh" "ks Tanh" "rd2d Tanh"; do set -- $w; run "" "" $1 $2; run "" "-DPDL_TANH_CACHE=1" $1 $2; done
run 1 "-DPDL_TANH_CACHE=1" heat Tanh
run "" "" heat Tanh; run "" "-DPDL_TANH_CACHE=1" heat Tanh
