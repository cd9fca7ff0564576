timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc $?" >> gpurun_out/r2r_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/r2r_smoke.log
timeout 400 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_c3.json 2> gpurun_out/r2r_bench_c3.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
