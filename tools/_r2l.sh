mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --windows 1 --no-e2e --no-cpu-baseline --no-library-baseline --no-also --no-graph --no-kernel-profile"
conv() { # rep -> raw csv (+ delete the report unless $2 = keep)
  ncu -i $1.ncu-rep --page raw --csv > $1_raw.csv 2>/dev/null; [ "$2" = keep ] || rm -f $1.ncu-rep; }
$B > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_syn-types.csv $B > gpurun_out/ncu1.log 2>&1; echo "launch list syn-types rc=$?"
$B > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:gru_persist" -s 6 -c 6 -o gpurun_out/r02_syn_gru $B > gpurun_out/ncu2.log 2>&1; echo "set full gru rc=$?"
ncu -i gpurun_out/r02_syn_gru.ncu-rep --page source --csv > gpurun_out/r02_syn_gru_source.csv 2>/dev/null; conv gpurun_out/r02_syn_gru
$B > gpurun_out/plain2b.log 2>&1 && ncu --set full --clock-control none -k "regex:gemm_tc|softmax_ce|gather_pool|adam_flat|tok_scatter" -s 60 -c 40 -o gpurun_out/r02_syn_rest $B > gpurun_out/ncu2b.log 2>&1; echo "set full rest rc=$?"
conv gpurun_out/r02_syn_rest
W="$B --workload wd-articles"
$W > gpurun_out/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_wd-articles.csv $W > gpurun_out/ncu3.log 2>&1; echo "launch list wd-articles rc=$?"
$W > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none -k "regex:gru_cluster|softmax_ce|gather_pool|tok_scatter|gemm_tc" -s 40 -c 30 -o gpurun_out/r02_wda_full $W > gpurun_out/ncu4.log 2>&1; echo "set full wd-articles rc=$?"
conv gpurun_out/r02_wda_full
du -sh gpurun_out; ls -la gpurun_out | head -30
