timeout 900 python -m pytest tests/test_gpu_threads_devlists.py tests/test_zzz_gpu_progressive.py tests/test_gpu_flights.py -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r2g_tests.log
timeout 900 python bench.py --workload c3 --volume-n 2048 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_c3_2048.json 2> gpurun_out/r2g_bench_c3_2048.err
FLIGHT_DEPTHS=4 FLIGHT_ENVS="GXY_PRIM_T=1" timeout 300 python tools/flight_sweep.py 1 24 > gpurun_out/r2g_sweep.log 2>&1
