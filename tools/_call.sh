timeout 900 python -m pytest tests/test_zz_gpu_pathlines.py tests/test_sampler.py tests/test_gpu_flights.py -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r2o_tests.log
timeout 300 python tools/f_rows_bench.py 1000 128 > gpurun_out/r2o_f_rows.json 2> gpurun_out/r2o_f_rows.err
GXY_FUSED_CURVES=0 timeout 300 python tools/f_rows_bench.py 1000 128 > gpurun_out/r2o_f_rows_list.json 2> gpurun_out/r2o_f_rows_list.err
timeout 600 python bench.py --workload pl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_pl.json 2> gpurun_out/r2o_bench_pl.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2o_smoke.log
