mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --windows 1 --no-e2e --no-cpu-baseline --no-library-baseline --no-also --no-graph --no-kernel-profile"
conv() { ncu -i $1.ncu-rep --page raw --csv > $1_raw.csv 2>/dev/null; rm -f $1.ncu-rep; }
$B > gpurun_out/plain1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02b_launches_syn-types.csv $B > gpurun_out/ncu1.log 2>&1; echo "launch list syn-types rc=$?"
W="$B --workload wd-articles"
$W > gpurun_out/plain3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02b_launches_wd-articles.csv $W > gpurun_out/ncu3.log 2>&1; echo "launch list wd-articles rc=$?"
$W > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none -k "regex:softmax_ce|gather_pool_bwd|reparam_kl" -s 2 -c 4 -o gpurun_out/r02b_wda_hbm $W > gpurun_out/ncu4.log 2>&1; echo "set full wd-articles hbm rc=$?"
conv gpurun_out/r02b_wda_hbm
du -sh gpurun_out
