# round-2 evidence run (one B200): GPU tests, headline bench, ncu launch list of the bench command, full captures of the
# dominant kernels. Outputs under gpurun_out/ (summaries are copied to profiles/ afterwards).
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_final.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --library-chi 0 --no-cpu-baseline --no-extra-configs --e2e-states 16384 > gpurun_out/r02_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jac -s 2 -c 1 -o gpurun_out/r02_kjac_final -f \
  python tools/dev/dev_prof.py methane-gri30 262144 jac > gpurun_out/ncu_kjac_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rates -s 2 -c 1 -o gpurun_out/r02_krates_final -f \
  python tools/dev/dev_prof.py methane-gri30 262144 rhs > gpurun_out/ncu_krates_final.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_btddod_solve_inv|k_btddod_invert" -s 2 -c 2 -o gpurun_out/r02_bt_final -f \
  python tools/dev/dev_prof_bt.py 1 > gpurun_out/ncu_bt_final.log 2>&1
python tools/dev/dev_peak.py > gpurun_out/r02_fp64_peak.txt 2>&1; cat gpurun_out/r02_fp64_peak.txt
for F in 1 56; do python tools/dev/dev_bt.py $F 2>&1 | grep "F="; done > gpurun_out/r02_bt_timings.txt; cat gpurun_out/r02_bt_timings.txt
ls -la gpurun_out | tail -12
