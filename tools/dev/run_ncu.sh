export GB_JAC4=1
GB_JAC4_THREADS=${1:-512} GB_JAC4_PRODUCERS=${2:-4} timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_jac4 -s 2 -c 1 -o gpurun_out/r02_kjac4_b -f python tools/dev/dev_prof.py methane-gri30 14208 jac > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log
