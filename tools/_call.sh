TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
GXY_PEER_TIMEOUT_MS=8000 timeout 600 $TR --master-port 29511 tools/mp_parity.py > gpurun_out/r2p_parity2.log 2>&1; echo "parity rc $?" >> gpurun_out/r2p_parity2.log
GXY_VOLUME_FLIGHTS=1 GXY_PEER_TIMEOUT_MS=8000 timeout 600 $TR --master-port 29514 bench.py --gpus 2 --workload c3 --steps 16 --warmup 3 > gpurun_out/r2p_bench_c3_n2.json 2> gpurun_out/r2p_bench_c3_n2.err
GXY_VOLUME_FLIGHTS=1 GXY_PEER_TIMEOUT_MS=8000 timeout 600 $TR --master-port 29515 bench.py --gpus 2 --workload c4 --steps 16 --warmup 3 > gpurun_out/r2p_bench_c4_n2.json 2> gpurun_out/r2p_bench_c4_n2.err
