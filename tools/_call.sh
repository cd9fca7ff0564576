for t in 1 8; do GXY_LIB=$PWD/galaxy_b200/libgxy_b200_counters.so timeout 300 python tools/prof_frame.py $t 4 > gpurun_out/r2i_counters_tess$t.log 2>&1; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"primary_trace_kernel|fused_secondary_kernel" -s 6 -c 2 -o gpurun_out/r2i_trace python tools/prof_frame.py 1 5 > gpurun_out/r2i_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_launches_bench.log 2>&1
ls -la gpurun_out/ > gpurun_out/r2i_ls.txt
