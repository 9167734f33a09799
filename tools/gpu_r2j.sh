cd $GRAFT_REPO_ROOT
run() { echo "== $*"; env "$@" QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=3 DEVA_B200_TILE_VERBOSE=1 python tools/tile_probe.py 2>&1 | grep -E "blocks/SM|rep\": 2" | cut -c1-330; }
run A=1
run QB_LIB=devastator_b200/libdeva_b200_m5.so DEVA_B200_TILE_RCAP=960
run QB_LIB=devastator_b200/libdeva_b200_m6.so DEVA_B200_TILE_RCAP=736
run QB_LIB=devastator_b200/libdeva_b200_m6.so DEVA_B200_TILE_RCAP=800
run QB_LIB=devastator_b200/libdeva_b200_t512.so DEVA_B200_TILE_RCAP=2560
run QB_LIB=devastator_b200/libdeva_b200_t512m3.so DEVA_B200_TILE_RCAP=1792
run QB_LIB=devastator_b200/libdeva_b200_n256.so DEVA_B200_TILE_RCAP=1280
run QB_LIB=devastator_b200/libdeva_b200_n256m3.so DEVA_B200_TILE_RCAP=960
run QB_P=0.5
run QB_LIB=devastator_b200/libdeva_b200_m6.so DEVA_B200_TILE_RCAP=736 QB_P=0.5
run QB_LIB=devastator_b200/libdeva_b200_t512m3.so DEVA_B200_TILE_RCAP=1792 QB_P=0.5
