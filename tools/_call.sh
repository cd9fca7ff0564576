timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r2e_tests.log
ENVS="GXY_FETCH_P=12;GXY_FETCH_P=8,GXY_FETCH_S=8;GXY_FETCH_P=4,GXY_FETCH_S=4;GXY_FETCH_P=16,GXY_FETCH_S=16;GXY_FETCH_P=8,GXY_FETCH_S=16;GXY_FETCH_P=16,GXY_FETCH_S=8;GXY_FETCH_P=24,GXY_FETCH_S=24;GXY_FUSED_BLOCKS_PER_SM=4;GXY_FUSED_BLOCKS_PER_SM=6"
FLIGHT_DEPTHS=4 FLIGHT_ENVS="$ENVS" timeout 600 python tools/flight_sweep.py 1 24 > gpurun_out/r2e_sweep_base.log 2>&1
GXY_LIB=$PWD/galaxy_b200/libgxy_b200_conv1.so FLIGHT_DEPTHS=4 FLIGHT_ENVS="GXY_FETCH_P=12;GXY_FETCH_P=8,GXY_FETCH_S=8" timeout 600 python tools/flight_sweep.py 1 24 > gpurun_out/r2e_sweep_conv1.log 2>&1
GXY_LIB=$PWD/galaxy_b200/libgxy_b200_conv2.so FLIGHT_DEPTHS=4 FLIGHT_ENVS="GXY_FETCH_P=12;GXY_FETCH_P=8,GXY_FETCH_S=8" timeout 600 python tools/flight_sweep.py 1 24 > gpurun_out/r2e_sweep_conv2.log 2>&1
timeout 400 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2e_bench_c3.json 2> gpurun_out/r2e_bench_c3.err
timeout 400 python bench.py --workload c4 --steps 10 --warmup 3 > gpurun_out/r2e_bench_c4.json 2> gpurun_out/r2e_bench_c4.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
