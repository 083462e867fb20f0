set -x
python tools/dev/dev_stagger.py 262144 1000 2000 4000 6000 7800 12000 2>&1 | tail -12
export GRIFFON_B200_LIB=spitfire_b200/libgriffon_b200_tl.so
python tools/timeline.py methane-gri30 37888 2>&1 | head -14
GB_JAC_GRID=4 python tools/timeline.py methane-gri30 1024 2>&1 | head -14
GB_JAC_STAGGER=7800 python tools/timeline.py methane-gri30 37888 2>&1 | head -14
