export GB_JAC4=1
for cfg in "512 4" "512 2" "640 4" "768 4" "768 6" "896 4" "1024 4"; do set -- $cfg; echo "== threads $1 producers $2"; GB_JAC4_THREADS=$1 GB_JAC4_PRODUCERS=$2 timeout 100 python tools/dev/dev_perf.py jac 2>&1 | tail -1; done
