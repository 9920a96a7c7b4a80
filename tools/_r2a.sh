set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests/test_baseline_shapes_gpu.py -x -q -m gpu -s > gpurun_out/r2a_shapes.log 2>&1; echo "shapes rc=$?"
tail -30 gpurun_out/r2a_shapes.log
python -m pytest tests -x -q -m gpu --deselect tests/test_baseline_shapes_gpu.py > gpurun_out/r2a_gpu.log 2>&1; echo "gpu rc=$?"
tail -5 gpurun_out/r2a_gpu.log
python tools/exp_graph_events.py > gpurun_out/r2a_events.log 2>&1; echo "events rc=$?"; tail -5 gpurun_out/r2a_events.log
python tools/library_baseline.py > gpurun_out/r2a_library.jsonl 2> gpurun_out/r2a_library.err; echo "lib rc=$?"; cat gpurun_out/r2a_library.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/r2a_smoke_ncu.log 2>&1; echo "ncu smoke rc=$?"; tail -3 gpurun_out/r2a_smoke_ncu.log; grep -c gru_cluster gpurun_out/r2a_smoke_launches.csv
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/gru_cluster_check.py 128 20 3 6 > gpurun_out/r2a_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -15 gpurun_out/r2a_racecheck.log
