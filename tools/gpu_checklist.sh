#!/bin/bash
# One gpurun call that re-establishes the GPU evidence after a change (run from the repo root on a B200 box):
#   gpurun --timeout 900 -- 'bash tools/gpu_checklist.sh r02_a'
# Writes everything under gpurun_out/<tag>_*; copy what should be judged into profiles/.
# Order: fast checks first, so that a cut-off call still leaves the most important files.
tag=${1:-check}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
# 1. parity: the whole -m gpu suite (the files sort so that the longest-established ones run first)
timeout 300 python -m pytest tests -m gpu -x -q > $out/${tag}_tests.log 2>&1; echo "tests rc=$?" >> $out/${tag}_tests.log
# 2. smoke
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1
# 3. headline bench (C5) and the CPU arm
timeout 400 python bench.py --steps 100 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
# 4. the rows added last: PathLines (contract line), PathLines + Sampler quick figures (both Sampler modes)
timeout 200 python bench.py --workload pl --steps 50 --warmup 3 > $out/${tag}_bench_pl.json 2> $out/${tag}_bench_pl.err
timeout 60 python tools/f_rows_bench.py 1000 128 > $out/${tag}_f_rows.json 2> $out/${tag}_f_rows.err
# 5. volume workloads
timeout 200 python bench.py --workload c3 --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err
timeout 200 python bench.py --workload c4 --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c4.json 2> $out/${tag}_bench_c4.err
# 6. ncu launch list of two C5 frames and of two PathLines frames (shares of the step, not absolute times)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/${tag}_launches_pl.csv \
    python bench.py --workload pl --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_ncu_bench_pl.log 2>&1
tail -2 $out/${tag}_tests.log; tail -1 $out/${tag}_smoke.log; cut -c1-200 $out/${tag}_bench.json; cut -c1-300 $out/${tag}_f_rows.json
