# closing run of round 2 (one B200): GPU tests, smoke, headline bench, full ncu capture of the final k_jac
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final3.log 2>&1; tail -2 gpurun_out/r02_pytest_gpu_final3.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final3.json 2> gpurun_out/r02_bench_final3.err; tail -c 300 gpurun_out/r02_bench_final3.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_jac -s 2 -c 1 -o gpurun_out/r02_kjac_final3 -f \
  python tools/dev/dev_prof.py methane-gri30 262144 jac > gpurun_out/ncu_kjac_final3.log 2>&1
ls -la gpurun_out | tail -5
