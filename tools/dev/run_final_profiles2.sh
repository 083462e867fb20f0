# round-2 closing evidence run (one B200): GPU tests, headline bench, full ncu captures of the twisted block-Thomas
# kernels, their timings. Outputs under gpurun_out/ (summaries are copied to profiles/ afterwards).
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final2.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_final2.log
python bench.py > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; tail -c 300 gpurun_out/r02_bench_final2.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_btddod_solve_inv|k_btddod_invert" -s 2 -c 2 -o gpurun_out/r02_bt_twisted -f \
  python tools/dev/dev_prof_bt.py 1 > gpurun_out/ncu_bt_twisted.log 2>&1
python tools/dev/dev_twist.py > gpurun_out/r02_twist_timings.txt 2>&1; cat gpurun_out/r02_twist_timings.txt
python tools/dev/dev_tick_stats.py 2>&1 | tail -4 > gpurun_out/r02_tick_stats.txt; cat gpurun_out/r02_tick_stats.txt
ls -la gpurun_out | tail -8
