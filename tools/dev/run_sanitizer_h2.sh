# compute-sanitizer over one small call of every kernel (h2-burke; the block-Thomas kernels are mechanism-independent)
out=gpurun_out/r02_compute_sanitizer_twist.txt; : > $out
for tool in memcheck racecheck synccheck; do
  echo "== $tool h2-burke" >> $out
  timeout 280 compute-sanitizer --tool $tool python tools/dev/dev_sanitize.py h2-burke 2>&1 | grep -v "^$" | tail -12 >> $out
done
tail -45 $out
