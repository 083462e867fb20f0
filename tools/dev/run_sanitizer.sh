# compute-sanitizer over one small call of every kernel (both mechanisms); output -> gpurun_out/r02_compute_sanitizer.txt
out=gpurun_out/r02_compute_sanitizer.txt; : > $out
for mech in h2-burke methane-gri30; do
  for tool in memcheck racecheck synccheck; do
    echo "== $tool $mech" >> $out
    timeout 900 compute-sanitizer --tool $tool python tools/dev/dev_sanitize.py $mech 2>&1 | grep -v "^$" | tail -12 >> $out
  done
done
tail -60 $out
