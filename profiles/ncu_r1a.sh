mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_r1a.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:front_kernel -s 3 -c 1 -o gpurun_out/prof_front_r1a $CMD > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sel_hist_kernel -s 18 -c 3 -o gpurun_out/prof_sel_r1a $CMD > gpurun_out/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:bin_hist_kernel|nuth_prepare" -s 12 -c 4 -o gpurun_out/prof_bin_r1a $CMD > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_l.log gpurun_out/ncu_f.log gpurun_out/ncu_s.log gpurun_out/ncu_b.log; ls -la gpurun_out
