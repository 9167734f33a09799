# final single-GPU evidence run: soak, bench (both arms), ncu launch list + full capture
set -x
cd $GRAFT_REPO_ROOT
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=30 python tools/tile_probe.py 2>&1 | grep -c MATCH
QB_P=0.5 QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=30 python tools/tile_probe.py 2>&1 | grep -c MATCH
QB_A=8388608 QB_P=0.5 QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=6 python tools/tile_probe.py 2>&1 | grep -c MATCH
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_n1_reference.json 2> gpurun_out/r2_bench_n1_reference.err; echo "ref rc=$?"
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_tile_probe.csv python tools/tile_probe.py > gpurun_out/r2_ncu_list.log 2>&1
QB_SKIP_SMALL=1 QB_ENGINES=tile QB_REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_step --launch-skip 120 --launch-count 8 -o gpurun_out/r2_k_tile_step python tools/tile_probe.py > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
python -c "import __graft_entry__ as g; g.smoke()"
