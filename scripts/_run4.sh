run() { PDL_NVCC_FLAGS="$1" python scripts/probe_one.py $2 $3 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$1\", /"; }
for w in "heat Tanh" "heat Rational" "rd2d Tanh" "ks Tanh" "burgers Rational"; do set -- $w; for fl in "" "-DPDL_WGRAD_ROLL_MT=1" "-DPDL_ADJ_FUSED=0"; do run "$fl" $1 $2; done; done
run "" heat Tanh; run "-DPDL_ADJ_FUSED=0" heat Tanh
