export ARK_GRU_CLUSTER_NO_COOP=1
CMD="python bench.py --workload wd-articles --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph"
$CMD > gpurun_out/plain_r1c_wda.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gru_cluster -s 6 -c 2 -f -o gpurun_out/prof_r1c_wda $CMD > gpurun_out/ncu_r1c_f.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/prof_r1c_wda.ncu-rep --page raw --csv > gpurun_out/prof_r1c_wda_raw.csv 2>/dev/null
ls -la gpurun_out/prof_r1c_wda.ncu-rep
CMD2="python bench.py --workload wd-articles --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD2 > gpurun_out/plain_r1c_wda_g.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1c_wd-articles.csv $CMD2 > gpurun_out/ncu_r1c_l.log 2>&1
echo "launches rc=$?"
tail -2 gpurun_out/plain_r1c_wda_g.log | cut -c1-300
grep -v "^$" gpurun_out/ncu_r1c_f.log | tail -4 | cut -c1-200
