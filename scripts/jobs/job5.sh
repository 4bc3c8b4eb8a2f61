timeout 300 ncu --set full --clock-control none --import-source on -k regex:mcs_resize_sep -s 18 -c 1 -o gpurun_out/prof_r2_resize python scripts/bench_prewarp.py > gpurun_out/ncu_resize.log 2>&1
tail -2 gpurun_out/ncu_resize.log | cut -c1-200
