export GB_JAC4=1
timeout 120 python tools/dev/dev_j4check.py methane-gri30 heptane-liu 2>&1 | grep -E "jac:|rhs:|bad|state" | head -40
GRIFFON_B200_LIB=spitfire_b200/libgriffon_b200_tl.so timeout 200 python tools/timeline4.py 2>&1 | tail -18
timeout 100 python tools/dev/dev_perf.py jac 2>&1 | tail -2
