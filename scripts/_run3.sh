run() { PDL_NVCC_FLAGS="$1" python scripts/probe_one.py $2 $3 1048576 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$1\", /"; }
for w in "burgers Tanh" "kdv Tanh" "ks Tanh" "rd2d Tanh"; do set -- $w; for fl in "" "-DPDL_WGRAD_BHOIST=1" "-DPDL_WGRAD_ROLL_MT=0" "-DPDL_WGRAD_NTB=5"; do run "$fl" $1 $2; done; done
for fl in "" "-DPDL_WGRAD_NTB=5" "-DPDL_WGRAD_NTB=2" "-DPDL_WGRAD_BHOIST=1"; do run "$fl" heat Tanh; done
for fl in "" "-DPDL_WGRAD_BHOIST=1" "-DPDL_WGRAD_NTB=5"; do run "$fl" heat Rational; run "$fl" ks Rational; done
